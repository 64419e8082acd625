"""Multi-GPU parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/mgpu_check.py

Every rank takes its column shard of one global batch; the sharded chain (peer-memory exchange, then
NCCL exchange) must reproduce the unsharded chain computed on the whole batch, and all ranks must
hold bit-identical statistics."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from parllel_b200 import ArrayDict  # noqa: E402
from parllel_b200.distributed import column_shard, shard_statistics  # noqa: E402
from parllel_b200.transforms import (ClipRewards, EstimateAdvantage, FusedBatchTransforms,  # noqa: E402
                                     NormalizeAdvantage, NormalizeObservations, NormalizeRewards)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T, B = 256, 64 * world * 4
    rng = np.random.default_rng(0)  # same global batch on every rank
    g = {"reward": (3 * rng.standard_normal((T, B))).astype(np.float32), "value": rng.standard_normal((T, B)).astype(np.float32),
         "done": rng.random((T, B)) < 0.02, "boot": rng.standard_normal(B).astype(np.float32)}

    def tree(sl):
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        return ArrayDict({"reward": d(g["reward"][:, sl]), "done": d(g["done"][:, sl]),
                          "agent_info": {"value": d(g["value"][:, sl])}, "bootstrap_value": d(g["boot"][sl]),
                          "advantage": torch.zeros(T, sl.stop - sl.start, device=dev),
                          "return_": torch.zeros(T, sl.stop - sl.start, device=dev),
                          "past_return": torch.zeros(T, sl.stop - sl.start, device=dev)})

    def chain(t):
        return [NormalizeRewards(t, 0.99), ClipRewards(-5, 5), EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(t)]

    # unsharded reference on every rank (no exchange: group of 1 is emulated by use of a private pipeline)
    os.environ["PLB_EXCHANGE"] = "p2p"
    whole = tree(slice(0, B))
    ref_tf = chain(whole)
    ref = FusedBatchTransforms(ref_tf, use_cuda_graph=False)
    ref._distributed = lambda: False
    ref(whole)
    sl = column_shard(B, world, rank)
    results = {}
    for mode in ("p2p", "nccl"):
        os.environ["PLB_EXCHANGE"] = mode
        mine = tree(sl)
        tfs = chain(mine)
        fused = FusedBatchTransforms(tfs, use_cuda_graph=(mode == "p2p"))
        for rep in range(3):  # eager, capture, replay
            mine["reward"].copy_(torch.from_numpy(np.ascontiguousarray(g["reward"][:, sl])).to(dev))
            tfs[0].return_statistics.load_state(0.0, 1.0, 1e-4)
            tfs[0]._carry_return = None
            tfs[0]._carry_done = None
            fused(mine)
        torch.cuda.synchronize()
        for k in ("reward", "past_return", "advantage", "return_"):
            got, want = mine[k].cpu().numpy(), whole[k][:, sl].cpu().numpy()
            bound = 2e-6 * np.maximum(np.abs(want), np.sqrt(np.mean(want.astype(np.float64) ** 2)))
            assert (np.abs(got - want) <= bound).all(), f"rank {rank} {mode} {k}: max err {np.abs(got - want).max()}"
        stats = torch.cat([tfs[0].return_statistics.mean_tensor, tfs[0].return_statistics.var_tensor,
                           tfs[0].return_statistics.count_tensor, fused._moments])
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        assert all(torch.equal(gathered[0], x) for x in gathered), f"{mode}: statistics differ between ranks"
        results[mode] = stats.clone()
    assert torch.equal(results["p2p"], results["nccl"]), "p2p and nccl exchanges disagree"

    # sharded NormalizeObservations (per-feature statistics, column layout)
    os.environ["PLB_EXCHANGE"] = "nccl"
    obs = (5 + 3 * rng.standard_normal((3, B, 6, 5))).astype(np.float32)
    full = torch.from_numpy(obs).to(dev)
    tf_full = NormalizeObservations({"observation": full}, obs_shape=(6, 5), initial_count=100)
    tf_full.normalize_batch(full)
    part = torch.from_numpy(np.ascontiguousarray(obs[:, sl])).to(dev)
    tf_part = shard_statistics(NormalizeObservations({"observation": part}, obs_shape=(6, 5), initial_count=100))
    tf_part.normalize_batch(part)
    assert torch.allclose(part, full[:, sl], rtol=1e-5, atol=1e-5), "sharded NormalizeObservations"
    assert np.allclose(tf_part.obs_statistics.var, tf_full.obs_statistics.var, rtol=1e-9)

    # timing of the two exchange flavours inside the fused step
    T2, B2 = 1024, 4096
    big = ArrayDict({"reward": torch.randn(T2, B2, device=dev), "done": torch.rand(T2, B2, device=dev) < 0.01,
                     "agent_info": {"value": torch.randn(T2, B2, device=dev)}, "bootstrap_value": torch.randn(B2, device=dev),
                     "advantage": torch.zeros(T2, B2, device=dev), "return_": torch.zeros(T2, B2, device=dev)})
    for mode in ("none", "p2p", "nccl"):
        os.environ["PLB_EXCHANGE"] = mode
        fz = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(big)])
        if mode == "none":  # every rank on its own shard statistics: isolates host/launch contention
            fz._distributed = lambda: False
        for _ in range(5):
            fz(big)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(500):
            fz(big)
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"sharded step (C1 per GPU, world {world}) with {mode} exchange: {e0.elapsed_time(e1) / 500 * 1e3:.1f} us")
        fz._graphs.clear(); fz._bound.clear()
    # latency of the stand-alone exchange kernel (all ranks ping each other back to back)
    from parllel_b200.distributed import PeerExchange
    ex = PeerExchange(3, None, dev)
    loc = torch.tensor([1.0 + rank, 2.0, 3.0], dtype=torch.float64, device=dev)
    out = torch.zeros(3, dtype=torch.float64, device=dev)
    for _ in range(10):
        ex.exchange(loc, out, 0, 1)
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(100):
            ex.exchange(loc, out, 0, 1)
    g.replay(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"stand-alone peer-memory exchange, world {world}: {e0.elapsed_time(e1) / 1000 * 1e3:.2f} us per exchange (kernel launch included)")
    if rank == 0:
        print("mgpu_check OK")
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
