"""Small driver that launches one hot-path kernel a few times (for ncu / timing sweeps).

    python tools/run_kernels.py gae [--T 1024 --B 4096 --algo chunked --reps 5]
    python tools/run_kernels.py norm_obs | gather | pipeline | moments
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from parllel_b200 import ArrayDict, BatchSpec  # noqa: E402
from parllel_b200.replays import ReplayBuffer  # noqa: E402
from parllel_b200.transforms import (EstimateAdvantage, NormalizeAdvantage,  # noqa: E402
                                     NormalizeObservations, NormalizeRewards)


def timed(fn, reps, flush=None, inner=1):
    """Per-launch time in us.  inner == 1: one event pair per launch after an L2 flush (event
    resolution on this box is ~2 us); inner > 1: `inner` back-to-back launches per event pair
    (callers rotate buffer sets larger than L2 through `i`)."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    graph = None
    if inner > 1:  # capture `inner` launches into one CUDA graph: no host overhead between kernels
        for k in range(inner):
            fn(k)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in range(inner):
                fn(k)
        graph.replay()
        torch.cuda.synchronize()
    for i in range(reps):
        if flush is not None and inner == 1:
            flush.zero_()
        ev[i][0].record()
        if graph is not None:
            graph.replay()
        else:
            fn(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) * 1e3 / inner for a, b in ev]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--T", type=int, default=1024)
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--algo", default="chunked")
    ap.add_argument("--lam", type=float, default=0.95)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--inner", type=int, default=1, help="launches per event pair (rotating buffer sets)")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--resident", type=int, default=1, help="resident advantage kernel: 0 never, 1 auto (multi-GPU only), 2 wherever it fits")
    ap.add_argument("--stages", type=int, default=0, help="cap on the input ring depth (0 = as many as fit)")
    ap.add_argument("--obs-variant", type=int, default=0)
    ap.add_argument("--debug", type=int, default=0)
    ap.add_argument("--obs-cols", type=int, default=16)
    ap.add_argument("--partials", type=int, default=1, help="gae_stats: publish per-CTA partials (as inside the fused step)")
    args = ap.parse_args()
    from parllel_b200 import _lib
    if args.rows:
        _lib.load().plb_set_tuning(0, args.rows)
    _lib.load().plb_set_tuning(12, args.resident)
    _lib.load().plb_set_tuning(9, args.stages)
    _lib.load().plb_set_tuning(3, args.obs_variant)
    _lib.load().plb_set_tuning(5, args.debug)
    _lib.load().plb_set_tuning(7, args.obs_cols)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)  # > L2
    T, B = args.T, args.B
    if args.what in ("gae", "normadv", "past", "gae_stats", "step", "chain"):
        n_sets = 1 if args.inner == 1 else max(2, int(np.ceil(2.5 * 126 * 2**20 / (17 * T * B))))
        trees = []
        for _ in range(n_sets):
            t = {"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < 0.01,
                 "agent_info": {"value": torch.randn(T, B, device=dev)}, "bootstrap_value": torch.randn(B, device=dev),
                 "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev),
                 "past_return": torch.empty(T, B, device=dev)}
            trees.append(ArrayDict(t))
        tree = trees[0]
        if args.what == "gae":
            tf = EstimateAdvantage(0.99, args.lam, algo=args.algo)
            nbytes = 17 * T * B
        elif args.what in ("gae_stats", "step", "chain"):
            from parllel_b200.transforms import ClipRewards, FusedBatchTransforms
            if args.what == "chain":
                lst = [NormalizeRewards(tree, 0.99), ClipRewards(-5, 5), EstimateAdvantage(0.99, args.lam), NormalizeAdvantage(tree)]
                nbytes = 25 * T * B
            else:
                lst = [EstimateAdvantage(0.99, args.lam), NormalizeAdvantage(tree)]
                nbytes = 17 * T * B
            fz = FusedBatchTransforms(lst, use_cuda_graph=False)
            ph = (_lib.PHASE_B | (_lib.PHASE_PARTIALS if args.partials else 0)) if args.what == "gae_stats" else _lib.PHASE_ALL
            tf = lambda tr: fz(tr, phases=ph)
        elif args.what == "normadv":
            for tr in trees:
                EstimateAdvantage(0.99, args.lam, algo=args.algo)(tr)
            tf = NormalizeAdvantage(tree)
            nbytes = 12 * T * B
        else:
            tf = NormalizeRewards(tree, 0.99, algo=args.algo)
            nbytes = 17 * T * B
        us = timed(lambda i: tf(trees[i % n_sets]), args.reps, flush, args.inner)
    elif args.what == "norm_obs":
        B, shape = 1024, (3, 84, 84)
        obs = torch.randn((4, B) + shape, device=dev) * 3 + 5
        tf = NormalizeObservations({"observation": obs}, obs_shape=shape, initial_count=10000)
        us = timed(lambda i: tf._step(obs[i % 4], None, dev), args.reps, flush, args.inner)
        nbytes = 8 * B * int(np.prod(shape))
    elif args.what == "gather":
        size_T, Bn = 62464, 16
        ring = ArrayDict({
            "observation": torch.randn(size_T + 2, Bn, 64, device=dev)[1:-1],
            "next_observation": torch.randn(size_T, Bn, 64, device=dev),
            "action": torch.randn(size_T, Bn, 8, device=dev),
            "reward": torch.randn(size_T, Bn, device=dev),
            "terminated": torch.rand(size_T, Bn, device=dev) < 0.01})
        rb = ReplayBuffer(ring, BatchSpec(128, Bn), size_T=size_T, replay_batch_size=args.n, oldest_n_samples_invalid=1)
        rb._full, rb._cursor = True, 384
        rb.seed(0)
        idx = [tuple(torch.from_numpy(a).to(dev) for a in rb.sample_indices()) for _ in range(4)]
        us = timed(lambda i: rb.gather(*idx[i % 4]), args.reps, flush, args.inner)
        nbytes = 1114 * args.n
    else:
        raise SystemExit(f"unknown kernel {args.what}")
    us = sorted(us)
    med = us[len(us) // 2]
    print(f"{args.what} T={T} B={B} algo={args.algo} rows={args.rows} resident={args.resident} debug={args.debug} stages={args.stages}: median {med:.2f} us  min {us[0]:.2f} us  "
          f"-> {nbytes / med / 1e3:.1f} GB/s (median), {nbytes / us[0] / 1e3:.1f} GB/s (best)  [{nbytes} algorithmic bytes]")


if __name__ == "__main__":
    main()
