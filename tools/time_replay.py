"""Replay timings on one GPU: index generation, gather and sample_batch() at the C3 shapes
(graph-timed: 8 calls per graph, best of 5 replays)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parllel_b200 import ArrayDict, BatchSpec, _lib  # noqa: E402
from parllel_b200.replays import ReplayBuffer  # noqa: E402


def graph_us(fn, n_inner=8, reps=5):
    for i in range(n_inner):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n_inner):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n_inner)
    return best


def main():
    dev = torch.device("cuda:0")
    size_T, Bn = 62464, 16
    ring = ArrayDict({
        "observation": torch.randn(size_T + 2, Bn, 64, device=dev)[1:-1],
        "next_observation": torch.randn(size_T, Bn, 64, device=dev),
        "action": torch.randn(size_T, Bn, 8, device=dev),
        "reward": torch.randn(size_T, Bn, device=dev),
        "terminated": torch.rand(size_T, Bn, device=dev) < 0.01})
    for variant in sys.argv[1:] or ["default"]:
        if "=" in variant:
            k, v = variant.split("=")
            _lib.load().plb_set_tuning(int(k), int(v))
        for n, k in ((256, 1), (16384, 1), (256, 512)):
            rb = ReplayBuffer(ring, BatchSpec(128, Bn), size_T=size_T, replay_batch_size=n, oldest_n_samples_invalid=1,
                              device_rng=True)
            rb._full, rb._cursor = True, 384
            rb.seed(0)
            idx = [rb.sample_indices_device(k) for _ in range(4)]
            us_g = graph_us(lambda i: rb.gather(idx[i % 4][0].reshape(-1), idx[i % 4][1].reshape(-1)))
            us_r = graph_us(lambda i: rb.sample_indices_device(k))
            us_s = graph_us(lambda i: rb.sample_batches(k) if k > 1 else rb.sample_batch())
            gbs = 1114 * n * k / us_g / 1e3
            print(f"[{variant}] n={n} k={k}: gather {us_g:.2f} us ({gbs:.0f} GB/s, {gbs / 6547.5:.2f} of peak)  "
                  f"indices {us_r:.2f} us  sample_batch(device_rng) {us_s:.2f} us", flush=True)


if __name__ == "__main__":
    main()
