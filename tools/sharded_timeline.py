"""Per-rank stamps of the sharded step (record-per-rank route) under torchrun: how long does each rank's normalise
kernel wait, and where does the rest of the step go?  (%globaltimer is per GPU: only differences within a rank are
meaningful.)   torchrun --nproc-per-node N tools/sharded_timeline.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parllel_b200 import ArrayDict, _lib
from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
T, B = 1024, 4096
trees = [ArrayDict({"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < 0.01,
                    "agent_info": {"value": torch.randn(T, B, device=dev)}, "bootstrap_value": torch.randn(B, device=dev),
                    "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev)}) for _ in range(6)]
for item in filter(None, os.environ.get("PLB_TUNING", "").split(",")):
    pass  # applied by _lib.load()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
lib.plb_debug_set_buffer(buf.data_ptr())  # set BEFORE capture: the pointer is a launch argument baked into the graphs
fz = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(trees[0])],
                          use_cuda_graph=os.environ.get("PLB_EAGER", "0") != "1")
for i in range(30):
    fz(trees[i % 6])
torch.cuda.synchronize(); dist.barrier()
rows = []
for rep in range(20):  # 20 bursts of 100 back-to-back graph replays; the stamps of each burst's LAST step are read
    for i in range(100):
        fz(trees[i % 6])
    torch.cuda.synchronize()
    t = buf.cpu().numpy().astype(np.float64)
    rows.append([(t[k] - t[0]) / 1e3 for k in range(7)])
    dist.barrier()
r = np.median(np.array(rows), axis=0)
names = ["scan start", "last CTA reached", "reduced", "pushed", "normalise start", "records seen", "normalise end"]
line = f"rank {rank}: " + "  ".join(f"{n} {v:6.2f}" for n, v in zip(names, r)) + f"   | wait {r[5] - r[4]:5.2f} us, launch gap {r[4] - r[3]:5.2f} us"
gathered = [None] * world
dist.all_gather_object(gathered, line)
if rank == 0:
    print(f"world {world}, PLB_TUNING={os.environ.get('PLB_TUNING', '')}: back-to-back graph replays, last step of 20 bursts (medians), us relative to the scan's start:")
    print("\n".join(gathered))
torch.cuda.synchronize(); dist.barrier()
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
h0 = time.perf_counter()
for i in range(600):
    fz(trees[i % 6])
h1 = time.perf_counter()
e1.record(); torch.cuda.synchronize()
host = torch.tensor([(h1 - h0) / 600 * 1e6], device=dev, dtype=torch.float64)
allh = [torch.zeros_like(host) for _ in range(world)]
dist.all_gather(allh, host)
if rank == 0:
    print(f"{'eager launches' if os.environ.get('PLB_EAGER', '0') == '1' else 'graph replay'}: {e0.elapsed_time(e1) / 600 * 1e3:.2f} us per step on the device; host time to ENQUEUE one step, per rank: "
          + " ".join(f"{float(h):.1f}" for h in allh) + " us")
lib.plb_debug_set_buffer(0)
dist.barrier()
os._exit(0)
