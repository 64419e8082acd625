"""Host time to ENQUEUE one C1 step through each launch path, and the device cadence it sustains (one GPU).
The sharded step re-synchronises all ranks every step, so a host that is only barely ahead of its GPU turns its
own jitter into everybody's wait.   python tools/host_cost.py"""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parllel_b200 import ArrayDict, _lib
from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage

dev = torch.device("cuda", 0)
lib = _lib.load()
T, B = (int(os.environ.get("PLB_T", 1024)), int(os.environ.get("PLB_B", 4096)))
N_SETS = 8
trees = [ArrayDict({"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < 0.01,
                    "agent_info": {"value": torch.randn(T, B, device=dev)}, "bootstrap_value": torch.randn(B, device=dev),
                    "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev)}) for _ in range(N_SETS)]
fz = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(trees[0])], launch="graph")
fd = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(trees[0])], launch="direct")
for i in range(4 * N_SETS):
    fz(trees[i % N_SETS])
    fd(trees[i % N_SETS])
torch.cuda.synchronize()
stream = torch.cuda.current_stream(dev)
raw_stream = torch._C._cuda_getCurrentRawStream(0)


def measure(name, call, n=3000):
    # (a) from an idle GPU: the first 200 calls never find the queue full -> pure enqueue cost
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        call(i)
    idle_us = (time.perf_counter() - t0) / 200 * 1e6
    torch.cuda.synchronize()
    # (b) steady state: per-call host time distribution and device cadence
    per = np.empty(n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(n):
        a = time.perf_counter_ns()
        call(i)
        per[i] = time.perf_counter_ns() - a
    e1.record(stream)
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) / n * 1e3
    per /= 1e3
    print(f"{name:34s}: enqueue from idle {idle_us:6.2f} us/call | steady: device {dev_us:6.2f} us/step, host per call "
          f"p50 {np.percentile(per, 50):6.2f}  p90 {np.percentile(per, 90):6.2f}  p99 {np.percentile(per, 99):6.2f}  max {per.max():8.2f} us", flush=True)


measure("fused(tree) [graph replay]", lambda i: fz(trees[i % N_SETS]))
measure("fused(tree) [direct, PDL on]", lambda i: fd(trees[i % N_SETS]))
lib.plb_set_tuning(18, 0)
measure("fused(tree) [direct, PDL off]", lambda i: fd(trees[i % N_SETS]))
lib.plb_set_tuning(18, 1)
# A/B inside one process (boxes of the pool differ by ~1 us per step): the normalise kernel's prologue
lib.plb_set_tuning(21, 1)
measure("direct, PDL on, serial partial loads (key 21 = 1)", lambda i: fd(trees[i % N_SETS]))
lib.plb_set_tuning(21, 0)
measure("direct, PDL on, one-round-trip prologue", lambda i: fd(trees[i % N_SETS]))
launch = lib.plb_graph_launch
execs = [fz._bound[(id(t), _lib.PHASE_ALL)][2][1] for t in trees]
measure("plb_graph_launch only", lambda i: launch(execs[i % N_SETS], raw_stream))
nop = lib.plb_sm_count
measure("ctypes no-op (plb_sm_count)", lambda i: nop(), n=1000)
