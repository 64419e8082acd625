"""Lean multi-rank timing of the sharded step (no parity, no CPU legs): every (route, launch mode) in one process group.
torchrun --nproc-per-node N tools/step_time.py [c1|c4] ...   -- median of 3 windows of 500 steps, max over ranks."""
import os, statistics, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parllel_b200 import ArrayDict, _lib
from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()


def trees_for(T, B, n):
    return [ArrayDict({"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < 0.01,
                       "agent_info": {"value": torch.randn(T, B, device=dev)}, "bootstrap_value": torch.randn(B, device=dev),
                       "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev)}) for _ in range(n)]


def timed(fz, trees, steps=500, windows=3):
    n = len(trees)
    for i in range(4 * n):
        fz(trees[i % n])
    out = []
    for _ in range(windows):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fz(trees[i % n])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t.item()))
    return statistics.median(out)


shapes = {"c1": (1024, 4096, 8), "c4": (256, 65536 // world, 8)}
for name in (sys.argv[1:] or ["c1", "c4"]):
    T, B, n = shapes[name]
    trees = trees_for(T, B, n)
    variants = [(3, "direct", 1024), (3, "direct", 512), (3, "direct", 256)] if world > 1 else [(0, "direct", 0), (0, "graph", 0)]
    for route, launch, threads in variants:
        lib.plb_set_tuning(17, route)
        if threads:
            lib.plb_set_tuning(20, threads)
        fz = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(trees[0])], launch=launch)
        us = timed(fz, trees)
        if rank == 0:
            print(f"world {world} {name} [T={T}, B={B}/rank] route {route} launch {launch} xch threads {threads}: {us:.2f} us/step", flush=True)
        del fz
    lib.plb_set_tuning(20, 256)
    del trees
lib.plb_set_tuning(17, 0)
if world > 1:
    dist.barrier()
os._exit(0)
