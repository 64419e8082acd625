"""The UNMODIFIED reference (balazsgyenes/parllel) timed on host cores: the CPU arm of bench.py.

`baseline/_ref/` holds `pip install --no-deps --target baseline/_ref /root/reference` (git-ignored, it
travels to the GPU box with the snapshot); `baseline/gymnasium_stub/` supplies the one import the
reference needs that this image lacks (type names + `gymnasium.utils.colorize`).  Nothing here is
product code: bench.py's `--impl reference` arm and its `cpu_baseline` legs are the only callers.

What is timed is the reference's own public API on its own leaf types (`ArrayDict[Array]`):
  * `EstimateAdvantage(...)(tree)` then `NormalizeAdvantage(tree)(tree)`   transforms/advantage.py:91-106,
    transforms/norm_advantage.py:30-56 -- numba (single thread: no `parallel=True` anywhere) + numpy;
  * `NormalizeObservations(...)(tree[t])`                                    transforms/norm_obs.py:57-76;
  * `ReplayBuffer.sample_batch()`                                            replays/replay.py:54-74 --
    numpy PCG64 draws + torch CPU advanced indexing (all intra-op threads).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUB_DIR = os.path.join(HERE, "gymnasium_stub")
_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "parllel"))


def load():
    """Import the installed reference (raises ImportError when baseline/_ref is absent)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError("baseline/_ref/parllel is missing (pip install --no-deps --target baseline/_ref /root/reference)")
    try:
        import gymnasium  # noqa: F401
    except ImportError:
        if STUB_DIR not in sys.path:
            sys.path.insert(0, STUB_DIR)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import parllel  # noqa: F401
    _loaded = parllel
    return parllel


def describe() -> dict:
    import numba
    import torch
    return {"numba": numba.__version__, "numba_threads_used": 1, "numpy": np.__version__,
            "torch_threads": torch.get_num_threads(), "host_cores": os.cpu_count()}


def _batch_tree(T, B):
    """Leaves as parllel/patterns.py allocates them (:155-163, :191-198, :341-346, :399-403)."""
    load()
    from parllel import Array, ArrayDict
    tree = ArrayDict()
    tree["reward"] = Array(batch_shape=(T, B), dtype=np.float32)
    tree["done"] = Array(batch_shape=(T, B), dtype=bool, padding=1)
    value = Array(batch_shape=(T, B), dtype=np.float32)
    tree["agent_info"] = ArrayDict({"value": value})
    tree["bootstrap_value"] = value[0].new_array(storage="local")
    tree["advantage"] = value.new_array(storage="local")
    tree["return_"] = value.new_array(storage="local")
    tree["past_return"] = tree["reward"].new_array(padding=1, storage="local")
    return tree


def _timed(fn, steps, warmup, budget_s):
    for _ in range(max(warmup, 1)):
        fn()
    times = []
    t_start = time.perf_counter()
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    return times


def time_advantage_step(batch: dict, gamma: float, lam: float, steps: int, warmup: int, budget_s: float,
                        full_chain: bool = False, return_outputs: bool = False):
    """One STEP of bench.py on the reference: EstimateAdvantage + NormalizeAdvantage (or the four-transform
    chain) on a [T, B] tree filled from `batch`.  Returns the list of per-step seconds (and, on request,
    the outputs of the last step so the caller can pin its oracle against them)."""
    load()
    from parllel.transforms import ClipRewards, EstimateAdvantage, NormalizeAdvantage, NormalizeRewards
    T, B = batch["reward"].shape
    tree = _batch_tree(T, B)
    tree["agent_info"]["value"][...] = batch["value"]
    tree["done"][...] = batch["done"]
    tree["bootstrap_value"][...] = batch["bootstrap_value"]
    tfs = [EstimateAdvantage(discount=gamma, gae_lambda=lam), NormalizeAdvantage(tree)]
    if full_chain:
        tfs = [NormalizeRewards(tree, discount=gamma), ClipRewards(-5.0, 5.0)] + tfs

    def step():
        t = tree
        t["reward"][...] = batch["reward"]  # the chain rescales rewards in place: refill (a 16 MB memcpy at C1,
        for tf in tfs:                       # < 1 % of the step; the GPU arm's inputs are not refilled either)
            t = tf(t)

    times = _timed(step, steps, warmup, budget_s)
    if return_outputs:
        return times, {k: np.asarray(tree[k]).copy() for k in ("advantage", "return_")}
    return times


def time_normalize_observations_step(B: int, obs_shape: tuple, steps: int, warmup: int, budget_s: float):
    load()
    from parllel import Array, ArrayDict
    from parllel.transforms import NormalizeObservations
    rng = np.random.default_rng(3)
    tree = ArrayDict()
    tree["observation"] = Array(batch_shape=(2, B), dtype=np.float32, feature_shape=obs_shape, padding=1)
    src = (5 + 3 * rng.standard_normal((B,) + tuple(obs_shape), dtype=np.float32)).astype(np.float32)
    tf = NormalizeObservations(tree, obs_shape=tuple(obs_shape), initial_count=10000)

    def step():
        tree["observation"][0] = src
        tf(tree[0])

    return _timed(step, steps, warmup, budget_s)


def time_replay_sample_batch(size_T: int, B: int, obs_dim: int, act_dim: int, n: int, steps: int, warmup: int,
                             budget_s: float, to_device=None):
    """ReplayBuffer.sample_batch() on a full ~1M-transition ring with SAC's leaves (torch/algos/sac.py:234-244);
    `to_device` adds the `.to(device)` the training script applies to every minibatch
    (examples/continuous_cartpole/train_sac.py batch_transform)."""
    load()
    import torch
    from parllel import Array, ArrayDict
    from parllel.replays import ReplayBuffer
    from parllel.types import BatchSpec
    rng = np.random.default_rng(4)
    T = 128
    tree = ArrayDict()
    obs = Array(batch_shape=(T, B), dtype=np.float32, feature_shape=(obs_dim,), full_size=size_T, padding=1)
    tree["observation"] = obs
    tree["next_observation"] = Array(batch_shape=(T, B), dtype=np.float32, feature_shape=(obs_dim,), full_size=size_T)
    tree["action"] = Array(batch_shape=(T, B), dtype=np.float32, feature_shape=(act_dim,), full_size=size_T)
    tree["reward"] = Array(batch_shape=(T, B), dtype=np.float32, full_size=size_T)
    tree["terminated"] = Array(batch_shape=(T, B), dtype=bool, full_size=size_T)
    for k in tree.keys():
        full = np.asarray(tree[k].full)
        if full.dtype == bool:
            full[...] = rng.random(full.shape, dtype=np.float32) < 0.01
        else:
            full[...] = rng.standard_normal(full.shape, dtype=np.float32)
    # as examples/continuous_cartpole/train_sac.py:118-123: the ring's `.full` views, torchified
    ring = ArrayDict({k: v.full for k, v in tree.items()}).to_ndarray().apply(torch.from_numpy)
    transform = (lambda t: t.to(device=to_device)) if to_device is not None else None
    rb = ReplayBuffer(tree=ring, sampler_batch_spec=BatchSpec(T, B), size_T=size_T, replay_batch_size=n,
                      oldest_n_samples_invalid=1, batch_transform=transform)
    rb._full, rb._cursor = True, 3 * T
    rb.seed(0)

    def step():
        out = rb.sample_batch()
        if to_device is not None:
            torch.cuda.synchronize()
        return out

    return _timed(step, steps, warmup, budget_s)
