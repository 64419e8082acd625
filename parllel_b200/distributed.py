"""Multi-GPU operation: the env axis ``B`` is sharded across the ranks of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The scans, the
elementwise passes and the replay gather shard with NO collective (independent env columns).
The only exchange step of the path is the statistics of the normalisers: each rank reduces its
shard to float64 ``(count, mean, M2)`` moments on the device, the moments are ALL-GATHERED
(24 bytes per rank for scalar statistics) and every rank merges them in rank order with the same
device kernel (``plb_merge_moments_f64``) -- an exact parallel (Chan) merge, bit-identical on
every rank, unlike the mean-of-variances all_reduce of parllel/torch/models.py:212-218.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._leaves import stream_ptr
from .transforms.norm_advantage import NormalizeAdvantage


def column_shard(B: int, world_size: int, rank: int) -> slice:
    """Contiguous block of env columns owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(B, world_size)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def all_gather_moments(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather a rank's float64 moments; returns ``[world, *local.shape]`` in rank order."""
    world = dist.get_world_size(group)
    flat = local.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat, group=group)  # flat in / flat out works for NCCL and gloo
    return out.reshape((world,) + tuple(local.shape))


def merge_gathered_moments(gathered: torch.Tensor, n_cols: int, layout: int, out: torch.Tensor) -> None:
    """Rank-ordered exact merge of gathered moments on the device (csrc/plb_stats.cu)."""
    rc = _lib.load().plb_merge_moments_f64(gathered.data_ptr(), gathered.shape[0], n_cols, layout, out.data_ptr(),
                                           stream_ptr(out.device))
    _lib.check(rc, "plb_merge_moments_f64")


class ShardedNormalizeAdvantage(NormalizeAdvantage):
    """``NormalizeAdvantage`` over a B-sharded batch: local moments -> all-gather -> merge ->
    local normalise.  Every rank normalises with the statistics of the WHOLE batch."""

    def __init__(self, sample_tree, group=None) -> None:
        super().__init__(sample_tree)
        self.group = group

    def _exchange(self, moments: torch.Tensor, inner: int) -> None:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        local = moments[: 3 * inner]
        gathered = all_gather_moments(local, self.group)
        merge_gathered_moments(gathered, inner, 0, local)


def shard_statistics(transform, group=None):
    """Switch a statistics-carrying transform (``NormalizeRewards``, ``NormalizeObservations``) to
    B-sharded operation: every update all-gathers the per-rank batch moments over ``group`` and merges
    them in rank order, so all ranks hold bit-identical running statistics of the WHOLE batch.
    Returns the transform for chaining."""
    for name in ("return_statistics", "obs_statistics"):
        stats = getattr(transform, name, None)
        if stats is not None:
            stats.sharded = True
            stats.group = group
    return transform
