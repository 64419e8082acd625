"""ctypes binding of ``libparllel_b200.so`` (the C ABI in ``include/parllel_b200.h``).

There is no CPU fallback: if the CUDA library is missing this module raises at
first use, loudly, instead of routing the hot path anywhere else.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libparllel_b200.so")

ALGO_AUTO, ALGO_SERIAL, ALGO_CHUNKED = 0, 1, 2
MOMENTS_F32_FLOW = 1
GATHER_MAX_LEAVES = 16
NAN = float("nan")


class PlbError(RuntimeError):
    """A libparllel_b200 entry point returned non-zero."""


class GatherLeaf(ctypes.Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("row_bytes", c_int64),
                ("stride_t_bytes", c_int64), ("stride_b_bytes", c_int64)]


class BatchPipeline(ctypes.Structure):
    """``plb_batch_pipeline`` (include/parllel_b200.h)."""
    _fields_ = [
        ("T", c_int64), ("B", c_int64),
        ("reward", c_void_p), ("ld_reward", c_int64),
        ("value", c_void_p), ("ld_value", c_int64),
        ("done", c_void_p), ("ld_done", c_int64),
        ("bootstrap_value", c_void_p),
        ("advantage", c_void_p), ("ld_advantage", c_int64),
        ("return_", c_void_p), ("ld_return", c_int64),
        ("past_return", c_void_p), ("ld_past_return", c_int64),
        ("previous_return", c_void_p), ("previous_done", c_void_p),
        ("valid", c_void_p), ("ld_valid", c_int64),
        ("normalize_rewards", c_int32), ("clip_rewards", c_int32),
        ("estimate_advantage", c_int32), ("normalize_advantage", c_int32),
        ("discount", c_double), ("gae_lambda", c_double),
        ("clip_lo", c_double), ("clip_hi", c_double),
        ("eps_reward", c_double), ("eps_advantage", c_double),
        ("return_mean", c_void_p), ("return_var", c_void_p), ("return_count", c_void_p),
        ("return_moments", c_void_p), ("advantage_moments", c_void_p),
        ("moments_flags", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("xch_peers", c_void_p), ("xch_rank", c_int32), ("xch_world", c_int32),
        ("carry_return_out", c_void_p), ("carry_done_out", c_void_p),
        ("gx_peers", c_void_p), ("gx_rank", c_int32), ("gx_world", c_int32), ("gx_slots", c_int32),
        ("gx_self", c_void_p),
        ("plan", c_int32),
    ]


PHASE_A, PHASE_B, PHASE_C, PHASE_ALL = 1, 2, 4, 7
PHASE_PARTIALS = 8  # modifier: B and C issued by separate calls, statistics travel as per-CTA partials
PLAN_AUTO, PLAN_CHUNKED, PLAN_RESIDENT = -1, 1, 2

_SIGNATURES = {
    "plb_version": ([], c_int),
    "plb_last_error": ([], ctypes.c_char_p),
    "plb_sm_count": ([], c_int),
    "plb_graph_launch": ([c_void_p, c_void_p], c_int),
    "plb_launch_counters": ([ctypes.POINTER(c_int64)], None),
    "plb_set_tuning": ([c_int, c_int], c_int),
    "plb_debug_set_buffer": ([c_void_p], None),
    "plb_compute_gae_advantage_f32": (
        [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
         c_int64, c_int64, c_int64, c_double, c_double, c_int, c_void_p], c_int),
    "plb_compute_discount_return_f32": (
        [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
         c_int64, c_int64, c_int64, c_double, c_int, c_void_p], c_int),
    "plb_compute_past_discount_return_f32": (
        [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
         c_double, c_int, c_void_p], c_int),
    "plb_batch_moments_workspace_bytes": ([], c_size_t),
    "plb_batch_moments_f32": (
        [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p], c_int),
    "plb_residual_moments_f32": (
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_size_t,
         c_void_p], c_int),
    "plb_valid_prefix_u8": ([c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p], c_int),
    "plb_column_moments_workspace_bytes": ([c_int64, c_int64], c_size_t),
    "plb_column_moments_f32": (
        [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
         c_void_p], c_int),
    "plb_merge_moments_f64": ([c_void_p, c_int32, c_int64, c_int, c_void_p, c_void_p], c_int),
    "plb_exchange_buffer_bytes": ([c_int32, c_int64], c_size_t),
    "plb_exchange_moments_p2p": ([c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int, c_int64, c_void_p, c_void_p], c_int),
    "plb_update_from_moments_f64": (
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p], c_int),
    "plb_scale_clip_f32": (
        [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_double, c_double, c_double, c_void_p], c_int),
    "plb_normalize_advantage_f32": (
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_double, c_void_p], c_int),
    "plb_normalize_columns_f32": (
        [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_double, c_void_p], c_int),
    "plb_normalize_columns_clip_f32": (
        [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_double, c_double, c_double, c_void_p], c_int),
    "plb_normalize_observations_workspace_bytes": ([c_int64, c_int64], c_size_t),
    "plb_normalize_observations_step_f32": (
        [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int,
         c_double, c_double, c_void_p, c_size_t, c_void_p], c_int),
    "plb_gather_tree": (
        [ctypes.POINTER(GatherLeaf), c_int32, c_void_p, c_void_p, c_int64, c_void_p], c_int),
    "plb_replay_sample_indices": (
        [c_void_p, c_int64, c_int64, c_int64, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p], c_int),
    "plb_batch_pipeline_workspace_bytes": ([], c_size_t),
    "plb_batch_pipeline_f32": ([ctypes.POINTER(BatchPipeline), c_int, c_void_p], c_int),
    "plb_batch_pipeline_plan": ([ctypes.POINTER(BatchPipeline)], c_int),
    "plb_grid_exchange_slots": ([], c_int),
    "plb_grid_exchange_bytes": ([c_int32, c_int32], c_size_t),
    "plb_grid_exchange_status": ([c_void_p, c_void_p], c_int),
}

_lib = None


def library_path() -> str:
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlbError(
            f"{LIB_PATH} is missing: build it with `python -m parllel_b200.build` "
            "(nvcc, sm_100a).  parllel_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = restype
    # experiments: PLB_TUNING="key=value,key=value" applies plb_set_tuning at load time
    for item in filter(None, os.environ.get("PLB_TUNING", "").split(",")):
        key, _, value = item.partition("=")
        if lib.plb_set_tuning(int(key), int(value)) != 0:
            raise PlbError(f"PLB_TUNING: bad item {item!r}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().plb_last_error()
        msg = msg.decode("utf-8", "replace") if msg else ""
        kind = "CUDA error" if rc > 0 else "library error"
        raise PlbError(f"{what or 'libparllel_b200'}: {kind} {rc}: {msg}")


def launch_counters() -> tuple[int, int]:
    """(kernels enqueued by the library so far, of which resident advantage kernels)."""
    out = (c_int64 * 2)()
    load().plb_launch_counters(out)
    return int(out[0]), int(out[1])


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)
