"""Stamps of the two-kernel grid-exchange consumer (normalize_advantage_gx_kernel) on one GPU (route forced)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parllel_b200 import ArrayDict, _lib
from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage
lib = _lib.load()
lib.plb_set_tuning(17, 2)
dev = torch.device("cuda:0")
T, B = 1024, 4096
trees = [ArrayDict({"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < 0.01,
                    "agent_info": {"value": torch.randn(T, B, device=dev)}, "bootstrap_value": torch.randn(B, device=dev),
                    "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev)}) for _ in range(6)]
fz = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(trees[0])], use_cuda_graph=False)
buf = torch.zeros(lib.plb_sm_count() * 8, dtype=torch.int64, device=dev)
for i in range(12):
    fz(trees[i % 6])
torch.cuda.synchronize()
lib.plb_debug_set_buffer(buf.data_ptr())
for i in range(6):
    fz(trees[i % 6])
torch.cuda.synchronize()
lib.plb_debug_set_buffer(0)
t = buf.cpu().numpy().reshape(-1, 8).astype(np.float64)
t0 = t[0, 0]
names0 = ["start", "before wait", "arrivals seen", "collected", "result stored", "after barrier", "retired", "end"]
print("collector CTA 0:", {n: round((t[0, k] - t0) / 1e3, 2) for k, n in enumerate(names0)})
names1 = ["start", "poll begins", "-", "-", "-", "result seen (after barrier)", "retired", "end"]
print("CTA 1:          ", {n: round((t[1, k] - t0) / 1e3, 2) for k, n in enumerate(names1) if n != "-"})
