"""Phase timeline of the resident advantage kernel (plb_debug_set_buffer): per-CTA %globaltimer stamps of the
LAST of a few back-to-back launches, relative to the earliest CTA start.

    python tools/resident_timeline.py [--T 1024 --B 4096] [key=value ...]     (plb_set_tuning pairs)
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parllel_b200 import ArrayDict, _lib  # noqa: E402
from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=1024)
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("tuning", nargs="*")
    args = ap.parse_args()
    lib = _lib.load()
    for item in args.tuning:
        k, v = item.split("=")
        lib.plb_set_tuning(int(k), int(v))
    dev = torch.device("cuda:0")
    T, B = args.T, args.B
    n_sets = max(2, int(np.ceil(2.5 * 126 * 2**20 / (17 * T * B))))
    trees = [ArrayDict({"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < 0.01,
                        "agent_info": {"value": torch.randn(T, B, device=dev)}, "bootstrap_value": torch.randn(B, device=dev),
                        "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev)})
             for _ in range(n_sets)]
    fz = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(trees[0])], use_cuda_graph=False)
    sms = lib.plb_sm_count()
    buf = torch.zeros(sms * 8, dtype=torch.int64, device=dev)
    lib.plb_debug_set_buffer(buf.data_ptr())
    for i in range(2 * n_sets):
        fz(trees[i % n_sets])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    buf.zero_()
    e0.record()
    for i in range(n_sets):
        fz(trees[i % n_sets])
    e1.record()
    torch.cuda.synchronize()
    lib.plb_debug_set_buffer(0)
    t = buf.cpu().numpy().reshape(sms, 8).astype(np.float64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    names = ["start", "scan done", "published", "collected", "stores issued", "stores drained"]
    print(f"T={T} B={B} tuning={args.tuning}: {len(t)} CTAs, {e0.elapsed_time(e1) * 1e3 / n_sets:.1f} us per eager launch")
    for k, name in enumerate(names):
        rel = (t[:, k] - t0) / 1e3
        print(f"  {name:15s} min {rel.min():7.2f}  median {np.median(rel):7.2f}  max {rel.max():7.2f} us")


if __name__ == "__main__":
    main()
