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
fz = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(trees[0])])
for i in range(30):
    fz(trees[i % 6])
torch.cuda.synchronize(); dist.barrier()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
rows = []
for rep in range(40):  # the debug pointer is a launch argument: eager launches only (graphs captured it as NULL)
    pass
fz2 = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(trees[0])], use_cuda_graph=False)
lib.plb_debug_set_buffer(buf.data_ptr())
for i in range(60):
    fz2(trees[i % 6])
    if i >= 20:
        torch.cuda.synchronize()
        t = buf.cpu().numpy().astype(np.float64)
        rows.append([(t[k] - t[0]) / 1e3 for k in range(7)])
        # keep the ranks loosely in step like a back-to-back run would
lib.plb_debug_set_buffer(0)
r = np.median(np.array(rows), axis=0)
names = ["scan start", "last CTA reached", "reduced", "pushed", "normalise start", "records seen", "normalise end"]
line = f"rank {rank}: " + "  ".join(f"{n} {v:6.2f}" for n, v in zip(names, r)) + f"   | wait {r[5] - r[4]:5.2f} us, launch gap {r[4] - r[3]:5.2f} us"
gathered = [None] * world
dist.all_gather_object(gathered, line)
if rank == 0:
    print(f"world {world}, eager launches (host-paced), medians of 40 steps, us relative to the scan's start:")
    print("\n".join(gathered))
# now back-to-back graph replays with timing, for reference
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(600):
    fz(trees[i % 6])
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"graph replay: {e0.elapsed_time(e1) / 600 * 1e3:.2f} us per step")
dist.barrier()
os._exit(0)
