"""bench.py -- headline benchmark of the parllel batch post-processing path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C1"): one STEP = one pass of the hot path over one synthetic
[T=1024, B=4096] float32 trajectory batch PER GPU (weak scaling; B is sharded, no data-path
collective): EstimateAdvantage (GAE, gamma=0.99, lambda=0.95) followed by NormalizeAdvantage
(whose (count, mean, M2) statistics are all-gathered and merged across ranks when N > 1).
`value`  = transitions/s, whole job, inputs resident in HBM, CUDA-event time, max over ranks.
`e2e`    = the same metric through the public Transform.__call__ API with HOST (pinned numpy)
           leaves: H2D of reward/value/done/bootstrap and D2H of advantage/return_ every step.
`roofline` is for the dominant kernel (the GAE scan): 17 algorithmic bytes / transition.
`cpu_baseline` / `--impl reference`: the reference algorithm (oracle port: C scans + numpy
reductions, oracle/) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_STEPS, B_ENVS = 1024, 4096          # C1, per GPU
GAMMA, LAMBDA, P_DONE = 0.99, 0.95, 0.01
GAE_BYTES_PER_TRANSITION = 17          # read reward 4 + value 4 + done 1, write advantage 4 + return_ 4
L2_BYTES = 126 * 2**20
METRIC = "transform_transitions_per_s"
UNIT = "transitions/s"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            return int(json.load(f)["traffic_bytes_per_launch"])
    except Exception:
        return None


def synth_batch(seed: int, T: int = T_STEPS, B: int = B_ENVS):
    rng = np.random.default_rng(seed)
    return dict(
        reward=rng.standard_normal((T, B), dtype=np.float32),
        value=rng.standard_normal((T, B), dtype=np.float32),
        done=rng.random((T, B), dtype=np.float32) < P_DONE,
        bootstrap_value=rng.standard_normal((B,), dtype=np.float32),
    )


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.05):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nvml = None

    def _sample(self):
        try:
            if self._nvml is not None:
                n = self._nvml
                self.samples.append(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM))
                self.max_mhz = n.nvmlDeviceGetMaxClockInfo(self._handle, n.NVML_CLOCK_SM)
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
                for name, bit in (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown),
                                  ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                                  ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown),
                                  ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)):
                    if mask & bit:
                        self.reasons.add(name)
            else:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                self.samples.append(int(parts[0]))
                self.max_mhz = int(parts[1])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:]):
                    if val.lower().startswith("active"):
                        self.reasons.add(name)
        except Exception:
            pass

    def __enter__(self):
        def loop():
            while not self._stop.is_set():
                self._sample()
                self._stop.wait(self.period)
        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._sample()
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self):
        return {"sm_mhz": int(statistics.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference_arm(steps: int, warmup: int, budget_s: float, threads: int, n_shards: int = 1):
    """The reference algorithm on host cores: oracle C scans (column blocks over a thread pool; the
    numba original is single-threaded) + numpy float32 reductions, same step as the GPU arm."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O

    O.lib()
    B_all = B_ENVS * n_shards  # the whole job's batch: every rank's shard, processed on the host
    batch = synth_batch(1, B=B_all)
    adv = np.empty_like(batch["value"])
    ret = np.empty_like(batch["value"])
    blocks = np.array_split(np.arange(B_all), max(threads, 1))
    pool = ThreadPoolExecutor(threads)

    def scan_block(cols):
        lo, hi = int(cols[0]), int(cols[-1]) + 1
        O.compute_gae_advantage(batch["reward"][:, lo:hi], batch["value"][:, lo:hi], batch["done"][:, lo:hi],
                                batch["bootstrap_value"][lo:hi], GAMMA, LAMBDA, adv[:, lo:hi], ret[:, lo:hi])

    def step():
        list(pool.map(scan_block, blocks))
        O.normalize_advantage(adv)

    for _ in range(max(warmup, 1)):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    pool.shutdown()
    per_step = sum(times) / len(times)
    return {"value": T_STEPS * B_all / per_step, "ms_per_step": per_step * 1e3, "steps": len(times)}


# ----------------------------------------------------------------------------- secondary kernels
def _graph_time_us(fn, n_inner, reps=5):
    """Average device time of one call of `fn(i)`: `n_inner` calls captured into a CUDA graph (no host
    gaps; CUDA events on this box tick at ~2 us), best of `reps` replays."""
    import torch
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
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n_inner)
    return best


def measure_extras(dev, peak):
    """The other BASELINE.json configs, timed on the device (inputs resident, rotating buffers > L2 where
    the working set is smaller than L2).  Reported next to the headline, not instead of it."""
    import torch

    from parllel_b200 import ArrayDict, BatchSpec
    from parllel_b200.replays import ReplayBuffer
    from parllel_b200.transforms import (ClipRewards, EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage,
                                         NormalizeObservations, NormalizeRewards)
    out = {}
    # C3: replay gather, ~1M-transition ring, obs 64 / act 8: 1114 algorithmic bytes per sample
    size_T, Bn = 62464, 16
    ring = ArrayDict({
        "observation": torch.randn(size_T + 2, Bn, 64, device=dev)[1:-1],
        "next_observation": torch.randn(size_T, Bn, 64, device=dev),
        "action": torch.randn(size_T, Bn, 8, device=dev),
        "reward": torch.randn(size_T, Bn, device=dev),
        "terminated": torch.rand(size_T, Bn, device=dev) < 0.01})
    for n, k in ((256, 1), (16384, 1), (256, 512)):
        rb = ReplayBuffer(ring, BatchSpec(128, Bn), size_T=size_T, replay_batch_size=n, oldest_n_samples_invalid=1,
                          device_rng=True)
        rb._full, rb._cursor = True, 384
        rb.seed(0)
        idx = [rb.sample_indices_device(k) for _ in range(4)]
        us = _graph_time_us(lambda i: rb.gather(idx[i % 4][0].reshape(-1), idx[i % 4][1].reshape(-1)), 8)
        gbs = 1114 * n * k / us / 1e3
        name = f"replay_gather_n{n}" + (f"_k{k}_fused" if k > 1 else "")
        out[name] = {"us_per_launch": us, "GB/s": gbs, "frac_of_measured_peak": gbs / peak, "samples_per_s": n * k / us * 1e6}
        if k == 1:
            us_rng = _graph_time_us(lambda i: rb.sample_batch(), 8)
            out[name]["us_sample_batch_incl_device_rng"] = us_rng
    # C2: one NormalizeObservations step, B=1024, obs 3x84x84: 8 bytes per observation element
    shape = (3, 84, 84)
    obs = torch.randn((4, 1024) + shape, device=dev) * 3 + 5
    tf = NormalizeObservations({"observation": obs}, obs_shape=shape, initial_count=10000)
    us = _graph_time_us(lambda i: tf._step(obs[i % 4], None, dev), 8)
    nbytes = 8 * 1024 * int(np.prod(shape))
    out["normalize_observations_step_B1024_3x84x84"] = {"us_per_step": us, "GB/s": nbytes / us / 1e3,
                                                        "frac_of_measured_peak": nbytes / us / 1e3 / peak,
                                                        "transitions_per_s": 1024 / us * 1e6}
    # full chain NormalizeRewards -> ClipRewards -> EstimateAdvantage -> NormalizeAdvantage at C1 and C4
    for T, B in ((T_STEPS, B_ENVS), (256, 65536)):
        n_sets = max(2, int(np.ceil(2.2 * L2_BYTES / (25 * T * B))))
        trees = []
        for _ in range(n_sets):
            trees.append(ArrayDict({"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < P_DONE,
                                    "agent_info": {"value": torch.randn(T, B, device=dev)},
                                    "bootstrap_value": torch.randn(B, device=dev),
                                    "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev),
                                    "past_return": torch.empty(T, B, device=dev)}))
        chain = FusedBatchTransforms([NormalizeRewards(trees[0], GAMMA), ClipRewards(-5, 5),
                                      EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(trees[0])], use_cuda_graph=False)
        us = _graph_time_us(lambda i: chain(trees[i % n_sets]), 2 * n_sets)
        out[f"full_chain_T{T}_B{B}"] = {"us_per_step": us, "transitions_per_s": T * B / us * 1e6,
                                        "GB/s_on_25B_per_transition": 25 * T * B / us / 1e3,
                                        "frac_of_measured_peak": 25 * T * B / us / 1e3 / peak}
        ea = EstimateAdvantage(GAMMA, LAMBDA)
        us = _graph_time_us(lambda i: ea(trees[i % n_sets]), 2 * n_sets)
        out[f"estimate_advantage_T{T}_B{B}"] = {"us_per_launch": us, "transitions_per_s": T * B / us * 1e6,
                                                "GB/s": 17 * T * B / us / 1e3, "frac_of_measured_peak": 17 * T * B / us / 1e3 / peak}
        del trees
    # C0 (T=128, B=16: the reference's own CartPole PPO batch): launch-latency-bound, so report wall-clock
    # microseconds per call of the public API (graph replay) instead of a roofline fraction
    T, B = 128, 16
    tree = ArrayDict({"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < P_DONE,
                      "agent_info": {"value": torch.randn(T, B, device=dev)}, "bootstrap_value": torch.randn(B, device=dev),
                      "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev),
                      "past_return": torch.empty(T, B, device=dev)})
    chain = FusedBatchTransforms([NormalizeRewards(tree, GAMMA), ClipRewards(-5, 5),
                                  EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(tree)])
    for _ in range(10):
        chain(tree)
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        chain(tree)
    torch.cuda.synchronize()
    out["full_chain_T128_B16"] = {"us_per_call_wall": (time.perf_counter() - t0) / n * 1e6, "bound": "launch latency",
                                  "api": "FusedBatchTransforms(4 transforms)(tree), CUDA-graph replay, device leaves"}
    return out


# ----------------------------------------------------------------------------- main
def _debug(msg):
    if os.environ.get("PLB_BENCH_DEBUG"):
        sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')} t={time.perf_counter():.2f}] {msg}\n")
        sys.stderr.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer e2e loop (0 = min(steps, 40))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary kernels (gather, observations, full chain)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"C1 per GPU: EstimateAdvantage(GAE gamma={GAMMA} lambda={LAMBDA}) + NormalizeAdvantage on synthetic "
                f"[T={T_STEPS},B={B_ENVS}] float32, done p={P_DONE}")
    config = {"workload": workload, "T": T_STEPS, "B_per_gpu": B_ENVS, "B_total": B_ENVS * max(world, 1),
              "sharding": "env axis B split across ranks; stats (count,mean,M2) all-gathered when N>1",
              "l2_policy": "inputs rotate over buffer sets totalling > 2x L2, so every step reads HBM"}
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        threads = min(cores, 32)
        res = cpu_reference_arm(args.steps, warmup, budget_s=60.0, threads=threads, n_shards=max(args.gpus, 1))
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": res["steps"], "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64 step / f32 store", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{res['steps']} steps over the whole job's batch [T={T_STEPS}, B={B_ENVS * max(args.gpus, 1)}] "
                                           f"(oracle C scans over {threads} threads + numpy reductions)"},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from parllel_b200 import ArrayDict, _lib
    from parllel_b200.host_pipeline import PinnedBatch
    from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # buffer sets: enough that consecutive steps never hit L2
    set_bytes = T_STEPS * B_ENVS * GAE_BYTES_PER_TRANSITION
    n_sets = max(3, int(np.ceil(2.2 * L2_BYTES / set_bytes)))
    sets = []
    for i in range(n_sets):
        b = synth_batch(1000 * rank + i + 1)
        t = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
        sets.append(ArrayDict({"reward": t["reward"], "done": t["done"], "agent_info": {"value": t["value"]},
                               "bootstrap_value": t["bootstrap_value"],
                               "advantage": torch.empty_like(t["value"]), "return_": torch.empty_like(t["value"])}))

    # the public API: the sampler's batch_transforms list, fused (one C call, CUDA-graph replay on 1 GPU;
    # phase-split with an NCCL all-gather of the (count, mean, M2) moments when B is sharded over ranks)
    fused = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(sets[0])])

    def step(i):
        fused(sets[i % n_sets])

    _debug("buffers ready")
    for i in range(max(warmup, 3 * n_sets)):  # every buffer set: eager pass, capture pass, replay pass
        step(i)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()

    _debug("warm-up done")
    stream = torch.cuda.current_stream(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for i in range(args.steps):
            step(i)
        ev1.record(stream)
        torch.cuda.synchronize()
        elapsed_ms = ev0.elapsed_time(ev1)
        _debug("timed region done")

        # dominant kernel alone, configured as inside the timed step (phase B publishing per-CTA partials = the
        # GAE scan with fused advantage moments):
        # one launch per buffer set captured into a CUDA graph, replayed until `steps` launches have run
        raw = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(sets[0])], use_cuda_graph=False)
        for i in range(n_sets):
            raw(sets[i], phases=_lib.PHASE_B | _lib.PHASE_PARTIALS)
        torch.cuda.synchronize()
        # >= 20 launches per graph so that the ~2.7 us device-side gap between two graph launches is not
        # charged to the kernel (CUDA event resolution is ~2 us as well)
        per_graph = n_sets * max(1, -(-20 // n_sets))
        kgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(kgraph):
            for i in range(per_graph):
                raw(sets[i % n_sets], phases=_lib.PHASE_B | _lib.PHASE_PARTIALS)
        kgraph.replay()
        torch.cuda.synchronize()
        n_replays = max(1, args.steps // per_graph)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(n_replays):
            kgraph.replay()
        k1.record(stream)
        torch.cuda.synchronize()
        kernel_launches = n_replays * per_graph
        gae_us = k0.elapsed_time(k1) * 1e3 / kernel_launches

        # keep the GPU under the same load for a few clock samples (a FIXED number of extra steps: with
        # N > 1 every step holds a collective, so all ranks must run the same count)
        for j in range(min(max(args.steps, 2000), 20000)):
            step(j)
        torch.cuda.synchronize()
    if distributed:
        dist.barrier()
        tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    transitions = T_STEPS * B_ENVS * world * args.steps
    value = transitions / (elapsed_ms * 1e-3)

    _debug("kernel loop + clock hold done")
    # ---- e2e: the same chain on HOST (pinned numpy) trees through the public host API
    # (HostBatchPipeline: H2D of batch i+1, kernels of batch i and D2H of batch i-1 overlap); every
    # step uploads its inputs and downloads its results inside the timed region
    from parllel_b200.host_pipeline import HostBatchPipeline
    e2e_steps = args.e2e_steps or min(args.steps, 60)
    host_sets = [PinnedBatch(synth_batch(7000 + 10 * rank + i)) for i in range(2)]
    e2e_fused = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(host_sets[0].tree)])
    pipe = HostBatchPipeline(e2e_fused, device=dev)
    for i in range(6):
        pipe(host_sets[i % 2].tree)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    prev = None
    for i in range(e2e_steps):
        h = pipe.submit(host_sets[i % 2].tree)
        if prev is not None:
            prev.wait()  # batch i-1 is complete in host memory before its buffers are reused
        prev = h
    prev.wait()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3  # host wall clock: results are in host memory at this point
    if distributed:
        tmax = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    e2e_value = T_STEPS * B_ENVS * world * e2e_steps / (e2e_ms * 1e-3)
    h2d = T_STEPS * B_ENVS * 9 + B_ENVS * 4   # reward, value, done, bootstrap_value
    d2h = T_STEPS * B_ENVS * 8                # advantage, return_

    def shutdown():
        """Leave cleanly: drop captured graphs (they pin NCCL work) before the process group; a watchdog
        forces the exit if NCCL teardown stalls, so the harness never waits on a finished benchmark."""
        sys.stdout.flush()
        if not distributed:
            return
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        try:
            for obj in (fused, e2e_fused):
                obj._graphs.clear()
                obj._bound.clear()
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
        finally:
            killer.cancel()

    if rank != 0:
        shutdown()
        return

    peak, peak_src = measured_peak()
    achieved = T_STEPS * B_ENVS * GAE_BYTES_PER_TRANSITION / (gae_us * 1e-6) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 scan carries / f32 storage", "data": "synthetic", "config": config,
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "api": "HostBatchPipeline(FusedBatchTransforms([EstimateAdvantage, NormalizeAdvantage])).submit(tree)/wait() on pinned numpy trees, 2 slots"},
        "gpu_launches": args.steps * 2,
        "kernels_per_step": ["scan_pipelined_kernel<GAE,+moments>", "normalize_advantage_partials_kernel"] if world == 1 else
                            ["scan_pipelined_kernel<GAE,+moments,+NVLink push>", "normalize_advantage_xch_kernel"],
        "roofline": {"bound": "hbm", "kernel": "plb::chunked::scan_pipelined_kernel<MODE=GAE, HAS_STATS> (EstimateAdvantage + fused advantage moments, per-CTA partials as in the step)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": measured_traffic(),
                     "traffic_source": "profiles/r1_traffic.json (ncu --set full of this command; outputs still dirty in L2 at kernel end)",
                     "algorithmic_bytes_per_launch": T_STEPS * B_ENVS * GAE_BYTES_PER_TRANSITION,
                     "avg_launch_us": gae_us, "launches_timed": kernel_launches,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
    }
    if not distributed and not args.no_extras:
        line["extras"] = measure_extras(dev, peak)
    if not args.no_cpu_baseline:
        res = cpu_reference_arm(steps=200, warmup=1, budget_s=12.0, threads=1)
        line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{res['steps']} full C1 steps, single thread like the numba original "
                                          f"(oracle C scans + numpy reductions); host has {cores} cores"}
    print(json.dumps(line))
    shutdown()


if __name__ == "__main__":
    main()
