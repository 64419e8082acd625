"""Pinned host <-> device copy bandwidth of the box (what bounds bench.py's `e2e` figure).
Measured on the round-1 B200 boxes: H2D 42.5 GB/s, D2H 52.7 GB/s, both directions at once ~46 GB/s each."""
import torch, time
x = torch.empty(64*2**20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(x, device="cuda")
y = torch.empty(64*2**20, dtype=torch.uint8).pin_memory(); e = torch.empty_like(y, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, n=20):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n):
        if h2d:
            with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
    torch.cuda.synchronize(); return 64*2**20*n/(time.perf_counter()-t)/1e9
print("H2D alone GB/s", run(True, False)); print("D2H alone GB/s", run(False, True)); print("both (each) GB/s", run(True, True))
