"""Multi-GPU parity worker (one process per GPU under torchrun; driven by tests/test_gpu_multirank.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/multirank_worker.py      (PLB_MULTIRANK_LOG=gpurun_out/multirank_wN.json keeps the report)

Every rank builds the SAME global batch from a seed, keeps its contiguous block of env columns
(parllel_b200.distributed.column_shard) and runs the sharded transforms.  Checked on every rank:
  * its shard of reward / past_return / advantage / return_ against the CPU ORACLE evaluated on the GLOBAL
    batch (oracle/: advantage.py:36-57, norm_rewards.py:86-116, norm_advantage.py:30-56), within
    1e-5 * max(|ref|, rms(ref));
  * the statistics every rank ends up with are bit-identical across ranks;
  * the NCCL route (PLB_EXCHANGE=nccl) agrees with the in-kernel routes within float32 rounding of the
    statistics;
  * uneven shards (some ranks TMA-describable, some not) agree on ONE exchange protocol and stay correct.
Cases: a C4-shaped shard per rank (T=256, B=65536 split over the ranks), the full four-transform chain over
three consecutive batches (eager / captured / replayed), lambda = 1, uneven B."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from parllel_b200 import ArrayDict, _lib  # noqa: E402
from parllel_b200.distributed import column_shard  # noqa: E402
from parllel_b200.transforms import (ClipRewards, EstimateAdvantage, FusedBatchTransforms,  # noqa: E402
                                     NormalizeAdvantage, NormalizeRewards)

GAMMA = 0.99


def synth(seed, T, B, scale=1.0):
    rng = np.random.default_rng(seed)
    return dict(reward=(scale * rng.standard_normal((T, B), dtype=np.float32)).astype(np.float32),
                value=rng.standard_normal((T, B), dtype=np.float32),
                done=rng.random((T, B), dtype=np.float32) < 0.01,
                bootstrap_value=rng.standard_normal((B,), dtype=np.float32))


def shard_tree(g, sl, dev, chain):
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    n = sl.stop - sl.start
    T = g["reward"].shape[0]
    leaves = {"reward": d(g["reward"][:, sl]), "done": d(g["done"][:, sl]),
              "agent_info": {"value": d(g["value"][:, sl])}, "bootstrap_value": d(g["bootstrap_value"][sl]),
              "advantage": torch.full((T, n), float("nan"), device=dev),
              "return_": torch.full((T, n), float("nan"), device=dev)}
    if chain:
        leaves["past_return"] = torch.full((T, n), float("nan"), device=dev)
    return ArrayDict(leaves)


def max_violation(got, ref):
    """max over elements of |got - ref| / (1e-5 * max(|ref|, rms(ref))): <= 1 passes."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    bound = 1e-5 * np.maximum(np.abs(ref), np.sqrt(np.mean(ref * ref)))
    return float(np.max(np.abs(got - ref) / bound)) if ref.size else 0.0


def identical_across_ranks(t, world):
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t.contiguous())
    return all(torch.equal(gathered[0], x) for x in gathered)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--report", default=os.environ.get("PLB_MULTIRANK_LOG", ""))
    args = ap.parse_args()
    args.log = args.report
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world, "cases": []}
    worst = 0.0

    def record(name, **kw):
        nonlocal worst
        viol = max(kw.get("violations", {}).values(), default=0.0)
        t = torch.tensor([viol], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = max(worst, float(t.item()))
        kw["max_violation_over_ranks"] = float(t.item())
        kw.pop("violations", None)
        report["cases"].append(dict(name=name, **kw))
        if rank == 0:
            print(f"[multirank w{world}] {name}: {kw}", flush=True)

    # ---- 1. C4-shaped step (EstimateAdvantage + NormalizeAdvantage): both exchange routes, lambda 0.95 and 1.0 ----
    T, B = 256, 65536
    routes = {3: "last-CTA reduction + record per rank (default)", 1: "resident kernel"}
    for lam, route in ((0.95, 3), (0.95, 1), (1.0, 3), (1.0, 1)):
        _lib.load().plb_set_tuning(17, route)
        g = synth(2, T, B)
        adv, ret = np.empty_like(g["value"]), np.empty_like(g["value"])
        fn = O.compute_discount_return if lam == 1.0 else O.compute_gae_advantage
        fn(g["reward"], g["value"], g["done"], g["bootstrap_value"], GAMMA, lam, adv, ret)
        O.normalize_advantage(adv)
        sl = column_shard(B, world, rank)
        os.environ["PLB_EXCHANGE"] = "p2p"
        tree = shard_tree(g, sl, dev, chain=False)
        fused = FusedBatchTransforms([EstimateAdvantage(GAMMA, lam), NormalizeAdvantage(tree)])
        before = _lib.launch_counters()
        for _ in range(4):  # eager, captured, replayed twice: same inputs, outputs overwritten each time
            tree["advantage"].fill_(float("nan"))
            fused(tree)
        torch.cuda.synchronize()
        after = _lib.launch_counters()
        v = {"advantage": max_violation(tree["advantage"].cpu().numpy(), adv[:, sl]),
             "return_": max_violation(tree["return_"].cpu().numpy(), ret[:, sl])}
        same = identical_across_ranks(fused._moments, world)
        assert same, "advantage statistics differ between ranks"
        status = fused._gx.status() if fused._gx is not None else -1
        assert status in (0, -1), "grid exchange timed out"
        # the NCCL route on the same shard
        os.environ["PLB_EXCHANGE"] = "nccl"
        tree2 = shard_tree(g, sl, dev, chain=False)
        fused2 = FusedBatchTransforms([EstimateAdvantage(GAMMA, lam), NormalizeAdvantage(tree2)], use_cuda_graph=False)
        fused2(tree2)
        torch.cuda.synchronize()
        v["advantage_nccl"] = max_violation(tree2["advantage"].cpu().numpy(), adv[:, sl])
        np.testing.assert_allclose(tree["advantage"].cpu().numpy(), tree2["advantage"].cpu().numpy(), rtol=3e-7, atol=3e-7)
        os.environ["PLB_EXCHANGE"] = "p2p"
        record(f"C4 step lambda={lam} route={routes[route]}", T=T, B_total=B, kernels_enqueued=after[0] - before[0],
               resident_launches=after[1] - before[1], stats_identical=same, violations=v)
        del tree, tree2, fused, fused2
    _lib.load().plb_set_tuning(17, 0)

    # ---- 2. the four-transform chain over three consecutive batches (state carried between them) -------
    T, B = 256, 1024 * world
    sl = column_shard(B, world, rank)
    tree = shard_tree(synth(100, T, B, 3.0), sl, dev, chain=True)
    tfs = [NormalizeRewards(tree, GAMMA), ClipRewards(-5, 5), EstimateAdvantage(GAMMA, 0.95), NormalizeAdvantage(tree)]
    fused = FusedBatchTransforms(tfs)
    stats = O.RunningMeanStd(())
    prev_ret, prev_done = np.zeros(B, np.float32), np.zeros(B, bool)
    v = {}
    for it in range(3):
        g = synth(100 + it, T, B, 3.0)
        for k, src in (("reward", "reward"), ("done", "done"), ("bootstrap_value", "bootstrap_value")):
            tree[k].copy_(torch.from_numpy(np.ascontiguousarray(g[src][..., sl])))
        tree["agent_info"]["value"].copy_(torch.from_numpy(np.ascontiguousarray(g["value"][:, sl])))
        fused(tree)
        torch.cuda.synchronize()
        r, past = g["reward"].copy(), np.empty_like(g["reward"])
        O.normalize_rewards(r, g["done"], past, prev_ret, prev_done, stats, GAMMA)
        prev_ret, prev_done = past[-1].copy(), g["done"][-1].copy()
        O.clip_rewards(r, -5, 5)
        adv, ret = np.empty_like(r), np.empty_like(r)
        O.compute_gae_advantage(r, g["value"], g["done"], g["bootstrap_value"], GAMMA, 0.95, adv, ret)
        O.normalize_advantage(adv)
        for k, ref in (("past_return", past), ("reward", r), ("return_", ret), ("advantage", adv)):
            v[f"b{it}_{k}"] = max_violation(tree[k].cpu().numpy(), ref[:, sl])
        st = tfs[0].return_statistics
        got = np.array([float(st.mean), float(st.var), float(st.count)])
        ref_st = np.array([float(stats.mean), float(stats.var), float(stats.count)])
        np.testing.assert_allclose(got, ref_st, rtol=2e-6)
        state = torch.cat([st.mean_tensor.reshape(-1), st.var_tensor.reshape(-1), st.count_tensor.reshape(-1), fused._moments])
        assert identical_across_ranks(state, world), "running statistics differ between ranks"
    record("four-transform chain, 3 batches", T=T, B_total=B, stats_identical=True, violations=v)
    del tree, fused

    # ---- 3. uneven shards: B = 16 * world + 1 gives rank 0 a width of 17 and the others 16: 16 is
    #         TMA-describable, 17 is not (done rows must be 16-byte multiples) -> the ranks must agree on ONE
    #         exchange protocol instead of each picking by its own shape ------------------------------------
    T, B = 96, 16 * world + 1
    g = synth(7, T, B)
    adv, ret = np.empty_like(g["value"]), np.empty_like(g["value"])
    O.compute_gae_advantage(g["reward"], g["value"], g["done"], g["bootstrap_value"], GAMMA, 0.95, adv, ret)
    O.normalize_advantage(adv)
    sl = column_shard(B, world, rank)
    tree = shard_tree(g, sl, dev, chain=False)
    fused = FusedBatchTransforms([EstimateAdvantage(GAMMA, 0.95), NormalizeAdvantage(tree)], use_cuda_graph=False)
    t0 = time.perf_counter()
    fused(tree)
    torch.cuda.synchronize()
    took = time.perf_counter() - t0
    assert took < 5.0, f"uneven shards stalled for {took:.1f} s (protocol mismatch?)"
    widths = [column_shard(B, world, r).stop - column_shard(B, world, r).start for r in range(world)]
    v = {"advantage": max_violation(tree["advantage"].cpu().numpy(), adv[:, sl]),
         "return_": max_violation(tree["return_"].cpu().numpy(), ret[:, sl])}
    assert identical_across_ranks(fused._moments, world)
    record("uneven shards", T=T, B_total=B, widths=widths, violations=v)

    ok = worst <= 1.0
    report["ok"] = ok
    report["worst_violation"] = worst
    if rank == 0:
        if args.log:
            os.makedirs(os.path.dirname(os.path.abspath(args.log)), exist_ok=True)
            with open(args.log, "w") as f:
                json.dump(report, f, indent=1)
        print(("MULTIRANK OK" if ok else "MULTIRANK FAILED") + f" world={world} worst_violation={worst:.3f}", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
