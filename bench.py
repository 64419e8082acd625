"""bench.py -- headline benchmark of the parllel batch post-processing path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--windows R]

Workload of the contract line (BASELINE.json configs[1], "C1"): one STEP = one pass of the hot path over
one synthetic [T=1024, B=4096] float32 trajectory batch PER GPU (weak scaling; B is sharded, no data-path
collective): EstimateAdvantage (GAE, gamma=0.99, lambda=0.95) followed by NormalizeAdvantage, whose
(count, mean, M2) statistics are exchanged between all CTAs of all ranks inside the kernel when N > 1.
`value`     transitions/s, whole job, inputs resident in HBM; CUDA-event time of exactly K steps, max over
            ranks, MEDIAN of R windows (every window is K steps between barrier + synchronize).
`e2e`       the same metric through the public Transform API on HOST (pinned numpy) trees: H2D of
            reward/value/done/bootstrap and D2H of advantage/return_ every step (pipelined throughput; the
            latency of one isolated batch is reported next to it).
`roofline`  the dominant kernel (17 algorithmic bytes/transition) + `secondary`: gather, observation step,
            past-return scan, full chain, C4 -- each with algorithmic bytes, launch time and fraction.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (baseline/_ref: numba + numpy, one thread as the
            reference runs them) on the box's host cores; the threaded C port of the oracle next to it.
`c4_sharded` BASELINE.json configs[4]: T=256, B=65536 split over the N ranks (strong scaling), with the
            step split into local work and exchange wait.
`parity_checked`: one un-timed step of every measured configuration compared with the CPU oracle on the
            GLOBAL batch (tolerance 1e-5 * max(|ref|, rms(ref))).
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
C4_T, C4_B = 256, 65536               # config 5: sharded over the ranks
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
    for name in ("r3_traffic.json", "r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
                return int(d["traffic_bytes_per_launch"]), f"profiles/{name}: {d.get('kernel', '')}"
        except Exception:
            continue
    return None, None


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


def bind_to_gpu_numa(index: int):
    """Pin this process (and the pinned buffers it allocates afterwards) to the CPUs NVML reports as local to
    GPU `index`: with several ranks per box, host threads and staging memory on the far socket halve the
    PCIe rate.  Returns the number of CPUs bound to, or None when the affinity is unknown."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------- CPU arms
def cpu_port_arm(steps: int, warmup: int, budget_s: float, threads: int, n_shards: int = 1):
    """The oracle's C restatement of the reference algorithm on host cores (column blocks over a thread
    pool + numpy float32 reductions): the `port` figure, faster than the numba original."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O

    O.lib()
    B_all = B_ENVS * n_shards
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
    per_step = statistics.median(times)
    return {"value": T_STEPS * B_all / per_step, "ms_per_step": per_step * 1e3, "steps": len(times)}


def cpu_reference_arm(steps: int, warmup: int, budget_s: float, n_shards: int = 1):
    """The UNMODIFIED reference (baseline/_ref) on the whole job's batch: EstimateAdvantage + NormalizeAdvantage
    through its own classes on its own Array leaves.  numba kernels and numpy reductions run on ONE thread
    (the reference has no parallel=True anywhere).  None when baseline/_ref is not installed."""
    from baseline import reference_arm as R
    if not R.available():
        return None
    B_all = B_ENVS * n_shards
    batch = synth_batch(1, B=B_all)
    times = R.time_advantage_step(batch, GAMMA, LAMBDA, steps, warmup, budget_s)
    per_step = statistics.median(times)
    return {"value": T_STEPS * B_all / per_step, "ms_per_step": per_step * 1e3, "steps": len(times),
            "env": R.describe()}


# ----------------------------------------------------------------------------- device timing helpers
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


def _violation(got, ref):
    """max |got - ref| / (1e-5 * max(|ref|, rms(ref))): <= 1 passes (tests/util.py's tolerance)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    bound = 1e-5 * np.maximum(np.abs(ref), np.sqrt(np.mean(ref * ref)))
    return float(np.max(np.abs(got - ref) / bound))


def _device_tree(batch, dev, sl=None, chain=False):
    import torch

    from parllel_b200 import ArrayDict
    cut = (lambda a: a) if sl is None else (lambda a: np.ascontiguousarray(a[..., sl]))
    t = {k: torch.from_numpy(cut(v)).to(dev) for k, v in batch.items()}
    leaves = {"reward": t["reward"], "done": t["done"], "agent_info": {"value": t["value"]},
              "bootstrap_value": t["bootstrap_value"],
              "advantage": torch.empty_like(t["value"]), "return_": torch.empty_like(t["value"])}
    if chain:
        leaves["past_return"] = torch.empty_like(t["value"])
    return ArrayDict(leaves)


def _oracle_step(batch):
    from oracle import oracle as O
    adv, ret = np.empty_like(batch["value"]), np.empty_like(batch["value"])
    O.compute_gae_advantage(batch["reward"], batch["value"], batch["done"], batch["bootstrap_value"], GAMMA, LAMBDA, adv, ret)
    O.normalize_advantage(adv)
    return adv, ret


def parity_check(make_global, T, B_total, rank, world, dev, dist):
    """One un-timed sharded step vs the oracle on the GLOBAL batch; returns the max violation over ranks."""
    import torch

    from parllel_b200.distributed import column_shard
    from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage
    g = make_global()
    adv, ret = _oracle_step(g)
    sl = column_shard(B_total, world, rank)
    tree = _device_tree(g, dev, sl)
    fz = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(tree)], use_cuda_graph=False)
    fz(tree)
    torch.cuda.synchronize()
    v = max(_violation(tree["advantage"].cpu().numpy(), adv[:, sl]), _violation(tree["return_"].cpu().numpy(), ret[:, sl]))
    if world > 1:
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        v = float(t.item())
    return v


LEAD_IN_STEPS = 3


def timed_windows(step, n_steps, n_windows, dev, dist, distributed):
    """`n_windows` windows of exactly `n_steps` steps, each between barrier + synchronize, CUDA-event time,
    max over ranks per window.  Returns the list of window times in ms.
    Between the synchronize and the first event every rank enqueues LEAD_IN_STEPS untimed steps: the events then
    bracket `n_steps` steps of a pipeline that is already flowing (the driver calls this with 20-step windows, where
    the host's first enqueue and -- with several ranks -- the skew with which the ranks leave the barrier, tens of
    microseconds once per window, would otherwise be charged to 20 steps; the lead-in steps' exchange re-aligns
    the ranks on the device)."""
    import torch
    stream = torch.cuda.current_stream(dev)
    out = []
    counter = 0
    for _ in range(n_windows):
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        for _ in range(LEAD_IN_STEPS):
            step(counter)
            counter += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n_steps):
            step(counter)
            counter += 1
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        out.append(ms)
    return out


# ----------------------------------------------------------------------------- secondary configurations (N = 1)
def measure_secondary(dev, peak, with_cpu):
    """The other BASELINE.json configurations on one GPU, inputs resident (rotating buffers > L2 where the
    working set is smaller than L2): device time per launch against algorithmic bytes, the reference's CPU
    time for the same call, and the end-to-end time of the public call where it has a host side."""
    import torch

    from parllel_b200 import ArrayDict, BatchSpec
    from parllel_b200.replays import ReplayBuffer
    from parllel_b200.transforms import (ClipRewards, EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage,
                                         NormalizeObservations, NormalizeRewards)
    rows, cpu, e2e = [], {}, {}

    def row(name, kernel, nbytes, us, **extra):
        gbs = nbytes / us / 1e3
        rows.append(dict(name=name, kernel=kernel, algorithmic_bytes_per_launch=int(nbytes), avg_launch_us=us,
                         achieved=gbs, unit="GB/s", frac=gbs / peak, **extra))

    # ---- C3: replay gather, ~1M-transition ring, obs 64 / act 8: 1114 algorithmic bytes per sample
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
        name = f"replay_gather_n{n}" + (f"_x{k}_fused" if k > 1 else "")
        row(name, "plb::gather_tree_kernel", 1114 * n * k, us, samples_per_s=n * k / us * 1e6,
            bound="launch latency" if n * k < 4096 else "hbm")
        us_idx = _graph_time_us(lambda i: rb.sample_indices_device(k), 8)
        us_all = _graph_time_us(lambda i: rb.sample_batches(k) if k > 1 else rb.sample_batch(), 8)
        e2e[name] = {"sample_batch_device_rng_us": us_all, "index_generation_us": us_idx,
                     "api": "ReplayBuffer(device_rng=True).sample_batch()" + ("es(k)" if k > 1 else "") + ": indices drawn on the device, no host traffic"}
        if k == 1:
            # default route: numpy draws on the host (bit-exact by construction) -> pinned staging -> H2D -> gather;
            # wall clock per call, minibatch complete on the device (synchronised) at the end of every call
            rb2 = ReplayBuffer(ring, BatchSpec(128, Bn), size_T=size_T, replay_batch_size=n, oldest_n_samples_invalid=1)
            rb2._full, rb2._cursor = True, 384
            rb2.seed(0)
            for _ in range(20):
                rb2.sample_batch()
            torch.cuda.synchronize()
            ts = []
            for _ in range(200):
                t0 = time.perf_counter()
                rb2.sample_batch()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            e2e[name]["sample_batch_host_rng_wall_us"] = statistics.median(ts) * 1e6
            e2e[name]["h2d_bytes_per_call"] = 16 * n
    # ---- C2: one NormalizeObservations step, B=1024, obs 3x84x84: 8 bytes per observation element
    shape = (3, 84, 84)
    obs = torch.randn((4, 1024) + shape, device=dev) * 3 + 5
    tf = NormalizeObservations({"observation": obs}, obs_shape=shape, initial_count=10000)
    us = _graph_time_us(lambda i: tf._step(obs[i % 4], None, dev), 8)
    row("normalize_observations_step_B1024_3x84x84", "plb::obs_lsu::obs_step_lsu_kernel", 8 * 1024 * int(np.prod(shape)), us,
        transitions_per_s=1024 / us * 1e6)
    del obs
    # ---- scans and chains at C1 and C4 on one GPU
    for T, B in ((T_STEPS, B_ENVS), (C4_T, C4_B)):
        n_sets = max(2, int(np.ceil(2.2 * L2_BYTES / (25 * T * B))))
        trees = []
        for _ in range(n_sets):
            trees.append(ArrayDict({"reward": torch.randn(T, B, device=dev), "done": torch.rand(T, B, device=dev) < P_DONE,
                                    "agent_info": {"value": torch.randn(T, B, device=dev)},
                                    "bootstrap_value": torch.randn(B, device=dev),
                                    "advantage": torch.empty(T, B, device=dev), "return_": torch.empty(T, B, device=dev),
                                    "past_return": torch.empty(T, B, device=dev)}))
        tag = f"T{T}_B{B}"
        chain = FusedBatchTransforms([NormalizeRewards(trees[0], GAMMA), ClipRewards(-5, 5),
                                      EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(trees[0])], use_cuda_graph=False)
        us = _graph_time_us(lambda i: chain(trees[i % n_sets]), 2 * n_sets)
        row(f"full_chain_{tag}", "past-return scan + advantage scan (+ normalise)", 25 * T * B, us, transitions_per_s=T * B / us * 1e6)
        ea = EstimateAdvantage(GAMMA, LAMBDA)
        us = _graph_time_us(lambda i: ea(trees[i % n_sets]), 2 * n_sets)
        row(f"estimate_advantage_{tag}", "plb::chunked::scan_pipelined_kernel<GAE>", 17 * T * B, us, transitions_per_s=T * B / us * 1e6)
        nr = NormalizeRewards(trees[0], GAMMA)
        us = _graph_time_us(lambda i: nr(trees[i % n_sets]), 2 * n_sets)
        row(f"normalize_rewards_{tag}", "plb::chunked::scan_pipelined_kernel<past return, +moments> + scale pass", 17 * T * B, us,
            transitions_per_s=T * B / us * 1e6)
        if (T, B) == (C4_T, C4_B):
            step = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(trees[0])], use_cuda_graph=False)
            us = _graph_time_us(lambda i: step(trees[i % n_sets]), 2 * n_sets)
            row(f"step_{tag}", "EstimateAdvantage + NormalizeAdvantage", 17 * T * B, us, transitions_per_s=T * B / us * 1e6)
        del trees
    # ---- C0 (T=128, B=16: the reference's own CartPole PPO batch): launch-latency-bound -> wall-clock us per call
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
    e2e["full_chain_T128_B16"] = {"us_per_call_wall": (time.perf_counter() - t0) / n * 1e6, "bound": "launch latency",
                                  "api": "FusedBatchTransforms(4 transforms)(tree), bound descriptor + programmatic dependent launch, device leaves"}
    # ---- the reference's own CPU code for the same calls (host cores of this box)
    if with_cpu:
        try:
            from baseline import reference_arm as R
            if R.available():
                med = lambda ts: statistics.median(ts)  # noqa: E731
                for n in (256, 16384):
                    cpu[f"replay_sample_batch_n{n}"] = {
                        "us_per_call": med(R.time_replay_sample_batch(size_T, Bn, 64, 8, n, 50, 3, 6.0)) * 1e6,
                        "us_per_call_incl_to_device": med(R.time_replay_sample_batch(size_T, Bn, 64, 8, n, 50, 3, 6.0, to_device=dev)) * 1e6,
                        "what": "reference ReplayBuffer.sample_batch(): numpy PCG64 draws + torch CPU indexing (all intra-op threads)"}
                cpu["normalize_observations_step_B1024_3x84x84"] = {
                    "us_per_call": med(R.time_normalize_observations_step(1024, shape, 5, 1, 8.0)) * 1e6,
                    "what": "reference NormalizeObservations(tree[t]): numpy, one thread"}
                cpu["full_chain_T1024_B4096"] = {
                    "us_per_call": med(R.time_advantage_step(synth_batch(1), GAMMA, LAMBDA, 8, 1, 8.0, full_chain=True)) * 1e6,
                    "what": "reference NormalizeRewards -> ClipRewards -> EstimateAdvantage -> NormalizeAdvantage: numba + numpy, one thread"}
                cpu["kind"] = "reference"
                cpu["env"] = R.describe()
        except Exception as err:  # the reference arm is context, never a reason to lose the GPU line
            cpu["error"] = repr(err)
    return rows, cpu, e2e


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
    ap.add_argument("--windows", type=int, default=5, help="timed windows of --steps steps each; the median is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer e2e loop (0 = min(steps, 60))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configurations (gather, observations, chains)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"C1 per GPU: EstimateAdvantage(GAE gamma={GAMMA} lambda={LAMBDA}) + NormalizeAdvantage on synthetic "
                f"[T={T_STEPS},B={B_ENVS}] float32, done p={P_DONE}")
    config = {"workload": workload, "T": T_STEPS, "B_per_gpu": B_ENVS, "B_total": B_ENVS * max(world, 1),
              "sharding": "env axis B split across ranks; (count,mean,M2) exchanged between all CTAs of all ranks inside the kernel when N>1",
              "l2_policy": "inputs rotate over buffer sets totalling > 2x L2, so every step reads HBM",
              "launch": "one plb_batch_pipeline_f32 call per step on the bound descriptor; kernels launched with programmatic dependent launch",
              "timing": f"median of {args.windows} windows of {args.steps} steps (CUDA events, barrier + synchronize around each, "
                        f"{LEAD_IN_STEPS} untimed lead-in steps between the synchronize and the first event, max over ranks)"}
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        n_shards = max(args.gpus, 1)
        res = cpu_reference_arm(args.steps, warmup, budget_s=60.0, n_shards=n_shards)
        kind, threads = "reference", 1
        sample = (f"{{steps}} steps over the whole job's batch [T={T_STEPS}, B={B_ENVS * n_shards}] through the unmodified reference "
                  "(baseline/_ref): EstimateAdvantage + NormalizeAdvantage on ArrayDict[Array]; numba + numpy, single thread "
                  "(the reference has no parallel=True)")
        if res is None:
            threads = min(cores, 32)
            res = cpu_port_arm(args.steps, warmup, budget_s=60.0, threads=threads, n_shards=n_shards)
            kind = "port"
            sample = (f"{{steps}} steps over the whole job's batch [T={T_STEPS}, B={B_ENVS * n_shards}] "
                      f"(baseline/_ref missing: oracle C scans over {threads} threads + numpy reductions)")
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": res["steps"], "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64 step / f32 store", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": threads, "kind": kind,
                                 "sample": sample.format(steps=res["steps"]), "host_cores": cores, **({"env": res["env"]} if "env" in res else {})},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from parllel_b200 import _lib
    from parllel_b200.distributed import column_shard
    from parllel_b200.host_pipeline import HostBatchPipeline, PinnedBatch
    from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    numa_cpus = None  # bound right before the host-buffer (e2e) leg: see below
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- parity first: one un-timed step of the measured configuration against the oracle on the global batch
    def c1_global():
        parts = [synth_batch(1000 * r + 1) for r in range(world)]
        return {k: np.concatenate([p[k] for p in parts], axis=-1) for k in parts[0]}
    parity = {"c1_max_violation": parity_check(c1_global, T_STEPS, B_ENVS * world, rank, world, dev, dist)}
    _debug("parity done")

    # ---- buffer sets: enough that consecutive steps never hit L2
    set_bytes = T_STEPS * B_ENVS * GAE_BYTES_PER_TRANSITION
    n_sets = max(3, int(np.ceil(2.2 * L2_BYTES / set_bytes)))
    sets = [_device_tree(synth_batch(1000 * rank + i + 1), dev) for i in range(n_sets)]

    # the public API: the sampler's batch_transforms list, fused (one C call per step on the bound descriptor)
    fused = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(sets[0])])

    def step(i):
        fused(sets[i % n_sets])

    before = _lib.launch_counters()
    step(0)  # eager pass: what the library enqueues for one step
    after = _lib.launch_counters()
    kernels_per_step, resident_per_step = after[0] - before[0], after[1] - before[1]
    for i in range(1, max(warmup, 3 * n_sets)):  # every buffer set: eager pass, capture pass, replay pass
        step(i)
    torch.cuda.synchronize()
    _debug("warm-up done")

    # clocks of the GPU whose rank prints the line; the other ranks do not poll NVML (every rank is coupled to every
    # other one through the statistics exchange, so a hiccup on any GPU is paid by all of them)
    with ClockSampler(local_rank, period_s=0.05 if rank == 0 else 3600.0) as clocks:
        windows_ms = timed_windows(step, args.steps, max(args.windows, 1), dev, dist, distributed)
        _debug("timed windows done")

        # dominant kernel alone, configured as inside the timed step: one launch per buffer set captured into a
        # CUDA graph (>= 20 launches per graph: the ~2.7 us gap between two graph launches and the ~2 us event
        # resolution are not charged to the kernel), replayed until `steps` launches have run
        raw = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(sets[0])], use_cuda_graph=False)
        if resident_per_step:  # the whole step is ONE kernel: scan + exchange + normalise
            kernel_name = ("plb::chunked::scan_pipelined_kernel<GAE, RESIDENT> (EstimateAdvantage + NormalizeAdvantage in one "
                           "cooperative launch: advantage parked in tensor memory, statistics exchanged between all CTAs)")
            launch = lambda i: raw(sets[i % n_sets])  # noqa: E731
        else:
            kernel_name = "plb::chunked::scan_pipelined_kernel<GAE, HAS_STATS> (EstimateAdvantage + fused advantage moments)"
            launch = lambda i: raw(sets[i % n_sets], phases=_lib.PHASE_B | _lib.PHASE_PARTIALS)  # noqa: E731
        for i in range(n_sets):
            launch(i)
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        per_graph = n_sets * max(1, -(-20 // n_sets))
        kgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(kgraph):
            for i in range(per_graph):
                launch(i)
        kgraph.replay()
        torch.cuda.synchronize()
        # at least 10 replays (>= 200 launches) whatever --steps is: the ~2 us event resolution and the one graph-launch
        # latency inside the bracket are then < 0.1 us per launch; one untimed replay leads in
        n_replays = max(10, args.steps // per_graph)
        stream = torch.cuda.current_stream(dev)
        kgraph.replay()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(n_replays):
            kgraph.replay()
        k1.record(stream)
        torch.cuda.synchronize()
        kernel_launches = n_replays * per_graph
        gae_us = k0.elapsed_time(k1) * 1e3 / kernel_launches
        del kgraph

        # keep the GPU under the same load for a few clock samples (a FIXED number of extra steps: with
        # N > 1 every step holds an exchange, so all ranks must run the same count)
        for j in range(min(max(args.steps, 2000), 20000)):
            step(j)
        torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    ms_per_step = statistics.median(windows_ms) / args.steps
    value = T_STEPS * B_ENVS * world / (ms_per_step * 1e-3)
    _debug("kernel loop + clock hold done")

    # ---- config 5 (C4): T=256, B=65536 split over the ranks (strong scaling), step split into local work / wait
    def c4_global():
        return synth_batch(2, C4_T, C4_B)
    parity["c4_max_violation"] = parity_check(c4_global, C4_T, C4_B, rank, world, dev, dist)
    sl = column_shard(C4_B, world, rank)
    B_loc = sl.stop - sl.start
    c4_sets_n = max(3, int(np.ceil(2.2 * L2_BYTES / (C4_T * B_loc * GAE_BYTES_PER_TRANSITION))))
    c4_sets = [_device_tree(synth_batch(50 + i, C4_T, C4_B), dev, sl) for i in range(min(c4_sets_n, 12))]
    c4_fused = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(c4_sets[0])])
    b0 = _lib.launch_counters()
    c4_fused(c4_sets[0])
    b1 = _lib.launch_counters()
    for i in range(1, 3 * len(c4_sets)):
        c4_fused(c4_sets[i % len(c4_sets)])
    c4_steps = max(200, min(args.steps, 1000))
    c4_ms = timed_windows(lambda i: c4_fused(c4_sets[i % len(c4_sets)]), c4_steps, max(args.windows, 1), dev, dist, distributed)
    c4_us = statistics.median(c4_ms) / c4_steps * 1e3
    c4 = {"workload": f"config 5: EstimateAdvantage + NormalizeAdvantage, T={C4_T}, B={C4_B} split over {world} rank(s) "
                      f"({B_loc} columns on rank 0), strong scaling",
          "us_per_step": c4_us, "value": C4_T * C4_B / (c4_us * 1e-6), "unit": UNIT,
          "kernels_per_step": b1[0] - b0[0], "resident_kernels_per_step": b1[1] - b0[1],
          "windows_us_per_step": [m / c4_steps * 1e3 for m in c4_ms], "steps_per_window": c4_steps,
          "frac_of_peak_whole_job": GAE_BYTES_PER_TRANSITION * C4_T * C4_B / (c4_us * 1e-6) / 1e9 / (measured_peak()[0] * world)}
    if distributed:
        # the same shard with the exchange confined to this GPU (a group of one): what the step costs without
        # waiting for the other ranks; the difference is the price of the cross-GPU exchange (latency + skew)
        local = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(c4_sets[0])])
        local._distributed = lambda: False
        for i in range(3 * len(c4_sets)):
            local(c4_sets[i % len(c4_sets)])
        loc_ms = timed_windows(lambda i: local(c4_sets[i % len(c4_sets)]), c4_steps, max(args.windows, 1), dev, dist, distributed)
        loc_us = statistics.median(loc_ms) / c4_steps * 1e3
        c4["local_only_us_per_step"] = loc_us
        c4["exchange_wait_us_per_step"] = c4_us - loc_us
    del c4_sets
    _debug("c4 done")

    # ---- e2e: the same chain on HOST (pinned numpy) trees through the public host API
    # (HostBatchPipeline: H2D of batch i+1, kernels of batch i and D2H of batch i-1 overlap); every
    # step uploads its inputs and downloads its results inside the timed region
    e2e_steps = args.e2e_steps or 60  # ~50 ms of transfers whatever --steps is: pipeline fill and drain stay < 2 %
    if world > 1:
        # host threads and pinned staging memory on the GPU's own socket (the device-resident legs above ran unbound,
        # like any launcher-started process)
        numa_cpus = bind_to_gpu_numa(local_rank)
    # three batches in flight: with two, batch i+1 can only be submitted once batch i-1 has been downloaded, and
    # kernels + download + host enqueue time (0.02 + 0.69 + ~0.15 ms) exceed the upload they should hide behind
    E2E_DEPTH = 3
    host_sets = [PinnedBatch(synth_batch(7000 + 10 * rank + i)) for i in range(E2E_DEPTH)]
    e2e_fused = FusedBatchTransforms([EstimateAdvantage(GAMMA, LAMBDA), NormalizeAdvantage(host_sets[0].tree)])
    pipe = HostBatchPipeline(e2e_fused, device=dev, n_slots=E2E_DEPTH)
    for i in range(3 * E2E_DEPTH):
        pipe(host_sets[i % E2E_DEPTH].tree)
    torch.cuda.synchronize()
    lat = []
    for i in range(10):  # one isolated batch at a time: upload, kernels, download, wait
        t0 = time.perf_counter()
        pipe(host_sets[i % E2E_DEPTH].tree)
        lat.append(time.perf_counter() - t0)
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    handles = []
    for i in range(e2e_steps):
        handles.append(pipe.submit(host_sets[i % E2E_DEPTH].tree))
        if i >= E2E_DEPTH - 1:
            # batch i-2 is complete in host memory before its buffers are reused (by batch i+1)
            handles[i - (E2E_DEPTH - 1)].wait()
    for h in handles[-(E2E_DEPTH - 1):]:
        h.wait()
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
        """Leave cleanly: drop captured graphs before the process group; a watchdog forces the exit if NCCL
        teardown stalls, so the harness never waits on a finished benchmark."""
        sys.stdout.flush()
        if not distributed:
            return
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        try:
            for obj in (fused, e2e_fused, c4_fused):
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
    traffic, traffic_src = measured_traffic()
    parity_ok = all(v <= 1.0 for v in parity.values())
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 scan carries / f32 storage", "data": "synthetic", "config": config,
        "windows_ms_per_step": [m / args.steps for m in windows_ms],
        "clocks": clocks.summary(),
        "parity_checked": parity_ok,
        "parity": dict(parity, tolerance="max |got-ref| / (1e-5*max(|ref|, rms(ref))) <= 1, vs the CPU oracle on the global batch"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "single_batch_latency_ms": statistics.median(lat) * 1e3,
                "per_rank_h2d_GBs": h2d / (e2e_ms / e2e_steps * 1e-3) / 1e9, "per_rank_d2h_GBs": d2h / (e2e_ms / e2e_steps * 1e-3) / 1e9,
                "cpus_bound_per_rank": numa_cpus,
                "note": "value is PIPELINED throughput (batch i+1 uploads while i computes and i-1 / i-2 download); an RL iteration "
                        "that holds one batch sees single_batch_latency_ms",
                "api": "HostBatchPipeline(FusedBatchTransforms([EstimateAdvantage, NormalizeAdvantage]), n_slots=3).submit(tree)/wait() on pinned numpy trees, three batches in flight"},
        "gpu_launches": args.steps * kernels_per_step,
        "kernels_per_step": kernels_per_step,
        "kernels_per_step_source": "plb_launch_counters() around one step (every timed step is one plb_batch_pipeline_f32 call on the cached descriptor; its kernels are launched with programmatic dependent launch)",
        "roofline": {"bound": "hbm", "kernel": kernel_name,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": T_STEPS * B_ENVS * GAE_BYTES_PER_TRANSITION,
                     "avg_launch_us": gae_us, "launches_timed": kernel_launches,
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     "step_frac": T_STEPS * B_ENVS * GAE_BYTES_PER_TRANSITION / (ms_per_step * 1e-3) / 1e9 / peak},
        "c4_sharded": c4,
    }
    if not distributed and not args.no_extras:
        rows, cpu2, e2e2 = measure_secondary(dev, peak, with_cpu=not args.no_cpu_baseline)
        line["roofline"]["secondary"] = rows
        line["e2e"]["secondary"] = e2e2
        if cpu2:
            line["cpu_baseline_secondary"] = cpu2
    if not args.no_cpu_baseline and not distributed:
        res = cpu_reference_arm(steps=30, warmup=1, budget_s=12.0)
        port = cpu_port_arm(steps=200, warmup=1, budget_s=6.0, threads=1)
        if res is not None:
            line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": 1, "kind": "reference",
                                    "sample": f"{res['steps']} full C1 steps through the unmodified reference (baseline/_ref: numba + numpy, "
                                              f"single thread as the reference runs them); host has {cores} cores",
                                    "ms_per_step": res["ms_per_step"], "env": res["env"],
                                    "port": {"value": port["value"], "ms_per_step": port["ms_per_step"], "cores": 1,
                                             "what": "oracle C restatement (oracle/plb_oracle.c) + numpy reductions, single thread"}}
        else:
            line["cpu_baseline"] = {"value": port["value"], "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{port['steps']} full C1 steps, single thread like the numba original "
                                              f"(oracle C scans + numpy reductions; baseline/_ref missing); host has {cores} cores"}
    print(json.dumps(line))
    shutdown()


if __name__ == "__main__":
    main()
