"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py

Writes small ``.npz`` fixtures next to this script.  The fixtures pin (a) the CPU
oracle (tests/test_oracle_golden.py, ``-m "not gpu"``) and (b) the CUDA path
(tests/test_gpu_*.py, ``-m gpu``) against the reference's own numpy/numba code:
parllel/transforms/{advantage,norm_rewards,norm_obs,norm_advantage,clip_rewards,
running_mean_std,multi_advantage}.py and parllel/replays/{replay,batched_dataloader}.py.
Every case is seeded; inputs are stored with the outputs so the GPU box never
needs the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
from parllel import Array, ArrayDict  # noqa: E402
from parllel.replays import BatchedDataLoader, ReplayBuffer  # noqa: E402
from parllel.transforms import (ClipRewards, EstimateAdvantage, EstimateMultiAgentAdvantage,  # noqa: E402
                                NormalizeAdvantage, NormalizeObservations, NormalizeRewards)
from parllel.types import BatchSpec  # noqa: E402


def save(name: str, **arrays) -> None:
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def make_batch_tree(T, B, with_valid=False, value_feat=()):
    """Leaves as parllel/patterns.py allocates them (:155-163 reward f32, :191-198 done
    bool padding=1, :341-346 advantage/return_ like value, :399-403 past_return padding=1)."""
    tree = ArrayDict()
    tree["reward"] = Array(batch_shape=(T, B), dtype=np.float32)
    tree["done"] = Array(batch_shape=(T, B), dtype=bool, padding=1)
    value = Array(batch_shape=(T, B), dtype=np.float32, feature_shape=value_feat)
    tree["agent_info"] = ArrayDict({"value": value})
    tree["bootstrap_value"] = value[0].new_array(storage="local")
    tree["advantage"] = value.new_array(storage="local")
    tree["return_"] = value.new_array(storage="local")
    tree["past_return"] = tree["reward"].new_array(padding=1, storage="local")
    if with_valid:
        tree["valid"] = tree["done"].new_array(storage="local")
    return tree


def fill_batch(tree, rng, T, B, p_done, value_feat=(), with_valid=False, reward_scale=1.0):
    reward = (rng.standard_normal((T, B)) * reward_scale).astype(np.float32)
    value = rng.standard_normal((T, B) + value_feat).astype(np.float32)
    done = rng.random((T, B)) < p_done
    boot = rng.standard_normal((B,) + value_feat).astype(np.float32)
    tree["reward"][...] = reward
    tree["agent_info"]["value"][...] = value
    tree["done"][...] = done
    tree["bootstrap_value"][...] = boot
    out = dict(reward=reward, value=value, done=done, bootstrap_value=boot)
    if with_valid:
        # valid mask as RecurrentSampler builds it (samplers/recurrent.py:141-142):
        # valid[t+1] = valid[t] & ~done[t]
        valid = np.ones((T, B), bool)
        for t in range(1, T):
            valid[t] = valid[t - 1] & ~done[t - 1]
        tree["valid"][...] = valid
        out["valid"] = valid
    return out


# ------------------------------------------------------------------ EstimateAdvantage
def gen_advantage():
    cases = [  # (T, B, p_done, lambda)
        (1, 1, 0.0, 0.95), (1, 7, 0.5, 1.0), (2, 5, 0.3, 0.95), (3, 16, 1.0, 0.95),
        (64, 96, 0.01, 0.95), (64, 96, 0.01, 1.0), (100, 33, 0.2, 0.9), (100, 33, 0.2, 1.0),
        (257, 130, 0.0, 0.97), (128, 16, 0.01, 0.95), (513, 64, 0.05, 0.95), (513, 64, 0.05, 1.0),
    ]
    out = {}
    for i, (T, B, p, lam) in enumerate(cases):
        rng = np.random.default_rng(1000 + i)
        tree = make_batch_tree(T, B)
        ins = fill_batch(tree, rng, T, B, p)
        tree = EstimateAdvantage(discount=0.99, gae_lambda=lam)(tree)
        out[f"{i}_meta"] = np.array([T, B, p, lam, 0.99], np.float64)
        for k, v in ins.items():
            out[f"{i}_{k}"] = v
        out[f"{i}_advantage"] = np.asarray(tree["advantage"]).copy()
        out[f"{i}_return_"] = np.asarray(tree["return_"]).copy()
    out["n_cases"] = np.array(len(cases))
    save("advantage", **out)


# ------------------------------------------- C0 pipeline: 3 rotating batches, 4 transforms
def gen_pipeline(name, T, B, p_done, n_batches, with_valid, seed, lam=0.95, clip=(-5.0, 5.0),
                 initial_count=None, reward_scale=3.0):
    rng = np.random.default_rng(seed)
    tree = make_batch_tree(T, B, with_valid=with_valid)
    transforms = [
        NormalizeRewards(tree, discount=0.99, initial_count=initial_count),
        ClipRewards(reward_min=clip[0], reward_max=clip[1]),
        EstimateAdvantage(discount=0.99, gae_lambda=lam),
        NormalizeAdvantage(tree),
    ]
    out = {"meta": np.array([T, B, p_done, n_batches, int(with_valid), lam, clip[0], clip[1],
                             -1.0 if initial_count is None else initial_count], np.float64)}
    for it in range(n_batches):
        tree.rotate()  # what the sampler does before every batch (samplers/basic.py:76)
        ins = fill_batch(tree, rng, T, B, p_done, with_valid=with_valid, reward_scale=reward_scale)
        for k, v in ins.items():
            out[f"b{it}_in_{k}"] = v
        out[f"b{it}_in_prev_past_return"] = np.asarray(tree["past_return"][-1]).copy()
        out[f"b{it}_in_prev_done"] = np.asarray(tree["done"][-1]).copy()
        for tf in transforms:
            tree = tf(tree)
        for k in ("reward", "past_return", "advantage", "return_"):
            out[f"b{it}_out_{k}"] = np.asarray(tree[k]).copy()
        st = transforms[0].return_statistics
        out[f"b{it}_out_stats"] = np.array([st.mean, st.var, st.count], np.float64)
    save(name, **out)


# ------------------------------------------------------------ NormalizeObservations steps
def gen_norm_obs(name, B, obs_shape, n_steps, with_valid, initial_count, seed):
    rng = np.random.default_rng(seed)
    tree = ArrayDict()
    tree["observation"] = Array(batch_shape=(n_steps, B), dtype=np.float32,
                                feature_shape=obs_shape, padding=1)
    if with_valid:
        tree["valid"] = Array(batch_shape=(n_steps, B), dtype=bool)
    tf = NormalizeObservations(tree, obs_shape=obs_shape, initial_count=initial_count)
    obs_in = (5.0 + 3.0 * rng.standard_normal((n_steps, B) + obs_shape)).astype(np.float32)
    tree["observation"][...] = obs_in
    out = {"meta": np.array([B, n_steps, int(with_valid),
                             -1.0 if initial_count is None else initial_count], np.float64),
           "obs_shape": np.array(obs_shape, np.int64), "in_observation": obs_in}
    if with_valid:
        valid = rng.random((n_steps, B)) < 0.7
        valid[2] = True
        tree["valid"][...] = valid
        out["in_valid"] = valid
    for t in range(n_steps):
        tf(tree[t])  # step transform on sample_tree[t] (samplers/basic.py:84-86)
        out[f"s{t}_mean"] = tf.obs_statistics.mean.copy()
        out[f"s{t}_var"] = tf.obs_statistics.var.copy()
        out[f"s{t}_count"] = np.array(tf.obs_statistics.count)
    out["out_observation"] = np.asarray(tree["observation"]).copy()
    save(name, **out)


# ------------------------------------------------------------------ multi-agent advantage
def gen_multi_advantage():
    out = {}
    cases = [(32, 12, 3, 0.95), (32, 12, 3, 1.0), (17, 5, 2, 0.9)]
    for i, (T, B, N, lam) in enumerate(cases):
        rng = np.random.default_rng(3000 + i)
        tree = make_batch_tree(T, B, value_feat=(N,))
        ins = fill_batch(tree, rng, T, B, 0.1, value_feat=(N,))
        tree["action"] = ArrayDict({f"a{k}": Array(batch_shape=(T, B), dtype=np.float32) for k in range(N)})
        tf = EstimateMultiAgentAdvantage(tree, discount=0.99, gae_lambda=lam)
        tree = tf(tree)
        adv = np.asarray(tree["advantage"]).copy()
        ret = np.asarray(tree["return_"]).copy()
        tree = NormalizeAdvantage(tree)(tree)  # per-agent axes (norm_advantage.py:44-49)
        out[f"{i}_meta"] = np.array([T, B, N, lam, 0.99], np.float64)
        for k, v in ins.items():
            out[f"{i}_{k}"] = v
        out[f"{i}_advantage"] = adv
        out[f"{i}_return_"] = ret
        out[f"{i}_norm_advantage"] = np.asarray(tree["advantage"]).copy()
    out["n_cases"] = np.array(len(cases))

    # single critic: scalar value, advantage gets a trailing singleton axis (multi_advantage.py:78-85)
    T, B = 24, 10
    rng = np.random.default_rng(3100)
    tree = make_batch_tree(T, B)
    ins = fill_batch(tree, rng, T, B, 0.1)
    tree["advantage"] = tree["agent_info"]["value"].new_array(feature_shape=(1,), storage="local")
    tree = EstimateMultiAgentAdvantage(tree, discount=0.99, gae_lambda=0.95)(tree)
    for k, v in ins.items():
        out[f"single_{k}"] = v
    out["single_advantage"] = np.asarray(tree["advantage"]).copy()
    out["single_return_"] = np.asarray(tree["return_"]).copy()

    # (the reference's markov-game branch calls np.stack on a generator, multi_advantage.py:118-121,
    #  which numpy >= 1.24 rejects, so no golden vector can be produced for it with this image)
    save("multi_advantage", **out)


# ------------------------------------------------------------------------- ReplayBuffer
def gen_replay():
    import torch

    T, B, size_T, obs_dim, act_dim = 8, 4, 64, 6, 2
    rng = np.random.default_rng(77)
    # leaves as torch/algos/sac.py:234-244 builds them: `.full` views of ring Arrays
    sample = ArrayDict()
    sample["observation"] = Array(batch_shape=(T, B), dtype=np.float32, feature_shape=(obs_dim,),
                                  padding=1, full_size=size_T)
    sample["next_observation"] = Array(batch_shape=(T, B), dtype=np.float32, feature_shape=(obs_dim,),
                                       full_size=size_T)
    sample["action"] = Array(batch_shape=(T, B), dtype=np.float32, feature_shape=(act_dim,), full_size=size_T)
    sample["reward"] = Array(batch_shape=(T, B), dtype=np.float32, full_size=size_T)
    sample["terminated"] = Array(batch_shape=(T, B), dtype=bool, full_size=size_T)
    ring = {}
    for k in sample:
        full = sample[k].full
        vals = rng.standard_normal(full.shape)
        vals = (vals > 1.0) if k == "terminated" else vals.astype(np.float32)
        full[...] = vals
        ring[k] = vals
    tree = ArrayDict({k: sample[k].full for k in sample})
    tree = tree.to_ndarray().apply(torch.from_numpy)  # train_sac.py:122-123
    # coordinate leaves reveal the sampled (T, B) indices bit-exactly
    tt, bb = np.meshgrid(np.arange(size_T), np.arange(B), indexing="ij")
    tree["t_index"] = torch.from_numpy(tt.copy())
    tree["b_index"] = torch.from_numpy(bb.copy())
    out = {f"ring_{k}": v for k, v in ring.items()}
    out["meta"] = np.array([T, B, size_T, obs_dim, act_dim], np.int64)
    scenarios = [  # (name, batch, newest_inv, oldest_inv, n_next_iteration, seed, n_draws)
        ("partial", 7, 0, 0, 3, 0, 3),
        ("partial_newest", 16, 1, 0, 5, 1, 2),
        ("full", 256, 0, 1, 11, 2, 3),  # cursor wraps: 11*8 = 88 -> 24, full
        ("full_both", 33, 2, 3, 9, 3, 4),
        ("pow2", 64, 0, 0, 8, 4, 2),  # exactly full, cursor 0
    ]
    for name, batch, newest, oldest, n_iter, seed, n_draws in scenarios:
        rb = ReplayBuffer(tree, BatchSpec(T, B), size_T=size_T, replay_batch_size=batch,
                          newest_n_samples_invalid=newest, oldest_n_samples_invalid=oldest)
        rb.seed(seed)
        for _ in range(n_iter):
            rb.next_iteration()
        out[f"{name}_meta"] = np.array([batch, newest, oldest, n_iter, seed, n_draws,
                                        rb._cursor, int(rb._full)], np.int64)
        for d in range(n_draws):
            b = rb.sample_batch()
            for k in b:
                out[f"{name}_d{d}_{k}"] = b[k].numpy().copy()
    out["scenarios"] = np.array([s[0] for s in scenarios])
    save("replay", **out)


# -------------------------------------------------------------------- BatchedDataLoader
def gen_dataloader():
    import torch

    T, B = 16, 6
    rng = np.random.default_rng(55)
    leaves = {
        "observation": rng.standard_normal((T, B, 5)).astype(np.float32),
        "advantage": rng.standard_normal((T, B)).astype(np.float32),
        "return_": rng.standard_normal((T, B)).astype(np.float32),
        "action": rng.integers(0, 4, (T, B)).astype(np.int64),
    }
    out = {f"in_{k}": v for k, v in leaves.items()}
    out["meta"] = np.array([T, B], np.int64)
    for name, n_batches, drop_last, shuffle in [("even", 4, True, True), ("ragged_drop", 5, True, True),
                                                ("ragged_keep", 5, False, True), ("noshuffle", 3, True, False)]:
        tree = ArrayDict({k: torch.from_numpy(v.copy()) for k, v in leaves.items()})
        tt, bb = np.meshgrid(np.arange(T), np.arange(B), indexing="ij")
        tree["t_index"] = torch.from_numpy(tt.copy())
        tree["b_index"] = torch.from_numpy(bb.copy())
        dl = BatchedDataLoader(tree, BatchSpec(T, B), n_batches=n_batches, drop_last=drop_last, shuffle=shuffle)
        dl.seed(9)
        n = 0
        for epoch in range(2):
            for batch in dl.batches():
                for k in batch:
                    out[f"{name}_{n}_{k}"] = batch[k].numpy().copy()
                n += 1
        out[f"{name}_meta"] = np.array([n_batches, int(drop_last), int(shuffle), n], np.int64)
    save("dataloader", **out)


def gen_metrics():
    """explained_variance (torch/utils.py:135-156), RunningMeanStdModel (torch/models.py:193-230, single
    process) and the recurrent sampler's valid recursion (samplers/recurrent.py:102-103,141-142)."""
    import torch
    from parllel.torch.models import RunningMeanStdModel
    from parllel.torch.utils import explained_variance

    rng = np.random.default_rng(21)
    out = {}
    for i, (n, noise) in enumerate([(5000, 0.3), (1237, 2.0), (64, 0.0)]):
        y_true = (2.0 + 1.5 * rng.standard_normal(n)).astype(np.float32)
        y_pred = (y_true + noise * rng.standard_normal(n)).astype(np.float32)
        out[f"ev{i}_true"], out[f"ev{i}_pred"] = y_true, y_pred
        out[f"ev{i}"] = np.float64(explained_variance(torch.from_numpy(y_pred), torch.from_numpy(y_true)))
    const = np.full(100, 3.0, np.float32)
    out["ev_const"] = np.float64(explained_variance(torch.from_numpy(const * 0.5), torch.from_numpy(const)))

    model = RunningMeanStdModel((5, 3))
    xs = (4.0 + 2.0 * rng.standard_normal((3, 6, 10, 5, 3))).astype(np.float32)
    for k in range(3):
        model.update(torch.from_numpy(xs[k]))
        out[f"rms{k}_mean"], out[f"rms{k}_var"] = model.mean.numpy().copy(), model.var.numpy().copy()
        out[f"rms{k}_count"] = model.count.numpy().copy()
    out["rms_x"] = xs

    # the sampler's statement, step by step, on a [T + 1, B] valid array (the last row is the padding row)
    T, B = 37, 29
    done = rng.random((T, B)) < 0.07
    valid = np.zeros((T + 1, B), np.bool_)
    valid[0] = True
    valid[1:] = False
    for t in range(T):
        valid[t + 1] = np.logical_and(valid[t], np.logical_not(done[t]))
    out["valid_done"], out["valid_mask"] = done, valid[:T]
    save("metrics", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:  # regenerate selected fixtures only: python gen_golden.py multi_advantage
        for name in sys.argv[1:]:
            globals()[f"gen_{name}"]()
        sys.exit(0)
    gen_advantage()
    gen_pipeline("pipeline_c0", T=128, B=16, p_done=0.01, n_batches=3, with_valid=False, seed=7)
    gen_pipeline("pipeline_valid", T=40, B=24, p_done=0.08, n_batches=3, with_valid=True, seed=8,
                 initial_count=1.0)
    gen_pipeline("pipeline_lambda1", T=33, B=20, p_done=0.05, n_batches=2, with_valid=False, seed=9,
                 lam=1.0, clip=(-1.0, 2.0))
    gen_norm_obs("norm_obs", B=64, obs_shape=(3, 8, 8), n_steps=5, with_valid=False,
                 initial_count=10000, seed=11)
    gen_norm_obs("norm_obs_valid", B=48, obs_shape=(10,), n_steps=4, with_valid=True,
                 initial_count=None, seed=12)
    gen_multi_advantage()
    gen_replay()
    gen_dataloader()
    gen_metrics()
