"""Fused execution of the sampler's batch-transform list.

The samplers run ``for transform in batch_transforms: sample_tree = transform(sample_tree)``
(parllel/samplers/basic.py:124-125).  ``FusedBatchTransforms`` is itself a ``Transform`` that
takes such a list -- the chain the examples build, in this order, every element optional:

    NormalizeRewards -> ClipRewards -> EstimateAdvantage -> NormalizeAdvantage

and runs it as ONE C call (``plb_batch_pipeline_f32``, csrc/plb_pipeline.cu): the reward
rescale/clip rides on the advantage scan's load path and the advantage moments on its store path,
so the batch is streamed 3 times instead of 6.  Results are those of calling the transforms one
after the other; the wrapped transform objects keep owning their state (``return_statistics``).
On device-resident trees the filled descriptor is bound to the tree object, so that a sampler-loop
call is one ctypes call whose kernels are launched with programmatic dependent launch
(``launch="direct"``, default), or the launch sequence is captured into a CUDA graph per set of leaf
addresses and replayed (``launch="graph"``).

Pass ``[FusedBatchTransforms(batch_transforms)]`` to the sampler instead of ``batch_transforms``.
"""
from __future__ import annotations

import ctypes
import os

import torch

from .. import _lib
from .._leaves import Staging, previous_row, rows_2d, stream_ptr
from ..arrays import DeviceArray
from ..tree import ArrayDict
from .advantage import EstimateAdvantage
from .clip_rewards import ClipRewards
from .norm_advantage import EPSILON as ADV_EPSILON
from .norm_advantage import NormalizeAdvantage
from .norm_rewards import EPSILON as REWARD_EPSILON
from .norm_rewards import NormalizeRewards
from .base import Transform

_ORDER = (NormalizeRewards, ClipRewards, EstimateAdvantage, NormalizeAdvantage)


class FusedBatchTransforms(Transform):
    def __init__(self, transforms, use_cuda_graph: bool = True, group=None, launch: str | None = None) -> None:
        # how a step on device-resident leaves reaches the GPU once its leaves are known:
        #   "direct": the cached descriptor goes straight to plb_batch_pipeline_f32 -- the kernels are launched with
        #             programmatic dependent launch, so consecutive kernels (and consecutive steps) overlap their
        #             launch latency and prologues with the predecessor's tail;
        #   "graph":  the launch sequence is captured into a CUDA graph per set of leaf addresses and replayed
        #             (a graph-to-graph boundary is a full serialisation: measured ~3 us per step slower at C1).
        launch = launch or os.environ.get("PLB_LAUNCH", "direct")
        if launch not in ("direct", "graph"):
            raise ValueError("launch must be 'direct' or 'graph'")
        self.launch = launch if use_cuda_graph else "eager"
        self.transforms = list(transforms)
        self.norm_rewards = self.clip = self.estimate = self.norm_adv = None
        slot = -1
        for tf in self.transforms:
            kind = next((i for i, cls in enumerate(_ORDER) if isinstance(tf, cls)), None)
            if kind is None or kind <= slot:
                raise ValueError(
                    "FusedBatchTransforms handles [NormalizeRewards, ClipRewards, EstimateAdvantage, "
                    f"NormalizeAdvantage] in that order (each optional); got {type(tf).__name__} out of place")
            slot = kind
            if kind == 0:
                self.norm_rewards = tf
            elif kind == 1:
                self.clip = tf
            elif kind == 2:
                self.estimate = tf
            else:
                self.norm_adv = tf
        if self.norm_adv is not None and self.norm_adv.multiagent:
            raise ValueError("multi-agent advantage is not handled by the fused chain")
        self.use_cuda_graph = use_cuda_graph
        self.group = group  # torch.distributed group for B-sharded operation (None = default / single GPU)
        self._ws = None
        self._moments = None
        self._graphs: dict = {}
        self._seen: set = set()
        # replay fast path: (id(tree), phases) -> (identity + address checks of its leaves, captured graph)
        self._bound: dict = {}
        self._gx = None          # grid-exchange buffers of the resident advantage kernel
        self._plans: dict = {}   # descriptor bytes -> plan bits agreed across the ranks

    # ------------------------------------------------------------------------------------
    def _distributed(self) -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _exchange(self, moments: torch.Tensor) -> None:
        from ..distributed import exchange_moments
        exchange_moments(moments, 1, 0, self.group)

    _LEAF_NAMES = ("reward", "done", "advantage", "return_", "past_return", "valid", "bootstrap_value")

    @staticmethod
    def _leaf_signature(leaf):
        """What a captured graph depends on for one device leaf: the memory it exposes right now
        (window rotation included) and its geometry.  None for host leaves."""
        if isinstance(leaf, torch.Tensor):
            return (leaf.data_ptr(), leaf.shape, leaf.stride()) if leaf.is_cuda else None
        if isinstance(leaf, DeviceArray):
            base = leaf.base
            return (base.data_ptr(), leaf._start, leaf._T, base.shape, leaf.padding)
        return None

    def _grid_exchange(self, dev):
        """Buffers through which every CTA of every rank trades its partial advantage moments inside the
        resident kernel; on one GPU a plain zeroed device buffer, on several a symmetric (peer-mapped) one."""
        if self._gx is None or self._gx.device != dev:
            from ..distributed import GridExchange
            self._gx = GridExchange(self.group, dev, local=not self._distributed())
        return self._gx

    def _agree_plan(self, lib, d, key: bytes) -> int:
        """Fused routes every rank can take: AND of the per-rank plans (uneven shards may differ in
        alignment), computed once per descriptor -- a collective, so all ranks must bind the same trees
        in the same order."""
        plan = self._plans.get(key)
        if plan is None:
            plan = lib.plb_batch_pipeline_plan(ctypes.byref(d))
            if plan < 0:
                _lib.check(plan, "plb_batch_pipeline_plan")
            if self._distributed():
                import torch.distributed as dist
                # bitwise AND as a MIN over one flag per route (NCCL has no BAND)
                bits = (_lib.PLAN_CHUNKED, _lib.PLAN_RESIDENT)
                t = torch.tensor([1 if plan & b else 0 for b in bits], dtype=torch.int32, device=self._ws.device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
                plan = sum(b for b, on in zip(bits, t.tolist()) if on)
            if len(self._plans) > 256:
                self._plans.clear()
            self._plans[key] = plan
        return plan

    @staticmethod
    def _address(leaf) -> int:
        """Where a device leaf's visible window starts right now (rotation of a DeviceArray included)."""
        if type(leaf) is torch.Tensor:
            return leaf.data_ptr()
        if isinstance(leaf, DeviceArray):
            return leaf.base.data_ptr() + leaf._start * leaf.base.stride(0) * leaf.base.element_size()
        return leaf.data_ptr()

    def _still_bound(self, sample_tree, checks) -> bool:
        """The replay fast path's guard: every leaf the graph was captured for is still THE object the tree
        holds under that name (identity: a replaced leaf, or a new dict that reuses a dead dict's id, cannot
        pass, because the binding keeps the captured leaves alive) and still exposes the same memory."""
        get = sample_tree._dict.get if type(sample_tree) is ArrayDict else sample_tree.get
        for name, sub, leaf, address in checks:
            cur = get(name)
            if sub is not None and cur is not None:
                cur = cur._dict.get(sub) if type(cur) is ArrayDict else cur.get(sub)
            if cur is not leaf or self._address(leaf) != address:
                return False
        return True

    def __call__(self, sample_tree, phases: int = _lib.PHASE_ALL):
        # fast path: this tree object is bound, with exactly these leaves at exactly these addresses, either to a filled
        # descriptor (one C call launches the kernels -- also inside somebody else's stream capture) or to a captured
        # graph (replayed, unless a capture is going on: no nested graphs).  A few microseconds of host time, which
        # matters: the sharded step re-synchronises all ranks every ~25 us, so the slowest host paces them
        if self._bound:
            entry = self._bound.get((id(sample_tree), phases))
            if entry is not None and self._still_bound(sample_tree, entry[0]):
                launch, handle, dev_index = entry[2]
                if entry[1] is None:  # direct: the cached descriptor, on the caller's current stream (capturable)
                    rc = launch(handle, phases, torch._C._cuda_getCurrentRawStream(dev_index))
                    if rc:
                        _lib.check(rc, "fused transforms")
                    return sample_tree
                if not torch._C._cuda_isCurrentStreamCapturing():
                    if launch is not None:  # cudaGraphLaunch through the C ABI on the caller's current stream
                        rc = launch(handle, torch._C._cuda_getCurrentRawStream(dev_index))
                        if rc:
                            _lib.check(rc, "plb_graph_launch")
                    else:
                        entry[1].replay()
                    return sample_tree
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        nr, est, na = self.norm_rewards, self.estimate, self.norm_adv
        lib = _lib.load()
        st = Staging(nr.return_statistics.device if nr is not None else None)
        d = _lib.BatchPipeline()
        keep = []  # tensors that must outlive the launch
        snapshots = []  # True where a row vector had to be COPIED to make it dense: such a call must not be bound or
        #                 captured (a replay would read the copy, not what the tree holds by then)

        def flat(t: torch.Tensor) -> torch.Tensor:
            f = t.reshape(-1).contiguous()
            snapshots.append(f.data_ptr() != t.data_ptr())
            return f

        reward_leaf = sample_tree["reward"] if (nr or self.clip or est) else None
        reward = st.view(reward_leaf, writeback=bool(nr or self.clip)) if reward_leaf is not None else None
        done = st.view(sample_tree["done"]) if (nr or est) else None
        ref = reward if reward is not None else st.view(sample_tree["advantage"], writeback=True)
        if ref.dim() != 2:
            raise ValueError("the fused chain needs [T, B] leaves")
        T, B = ref.shape
        d.T, d.B = T, B
        if reward is not None:
            d.reward, d.ld_reward, _ = rows_2d(reward)
        if done is not None:
            d.done, d.ld_done, _ = rows_2d(done)
        valid = None
        if (nr is not None and nr.only_valid) or (na is not None and na.only_valid):
            valid = st.view(sample_tree["valid"])
            d.valid, d.ld_valid, _ = rows_2d(valid)

        past = None
        if nr is not None:
            past_leaf = sample_tree["past_return"]
            past = st.view(past_leaf, writeback=True)
            d.past_return, d.ld_past_return, _ = rows_2d(past)
            prev_ret = previous_row(past_leaf, st)
            prev_done = previous_row(sample_tree["done"], st)
            if prev_ret is None:
                if nr._carry_return is None:
                    nr._carry_return = torch.zeros(B, dtype=torch.float32, device=st.device)
                prev_ret = nr._carry_return
            if prev_done is None:
                if nr._carry_done is None:
                    nr._carry_done = torch.zeros(B, dtype=torch.bool, device=st.device)
                prev_done = nr._carry_done
            prev_ret, prev_done = flat(prev_ret), flat(prev_done)
            keep += [prev_ret, prev_done]
            d.previous_return, d.previous_done = prev_ret.data_ptr(), prev_done.data_ptr()
            if nr._carry_return is not None and nr._carry_done is not None:
                # leaves without padding rows: phase A itself leaves past_return[T-1] / done[T-1] in the
                # carry buffers (aliasing the previous rows is allowed)
                d.carry_return_out, d.carry_done_out = nr._carry_return.data_ptr(), nr._carry_done.data_ptr()
            stats = nr.return_statistics
            d.return_mean, d.return_var = stats.mean_tensor.data_ptr(), stats.var_tensor.data_ptr()
            d.return_count = stats.count_tensor.data_ptr()
            d.normalize_rewards = 1
            d.discount = float(nr.discount)
        if self.clip is not None:
            d.clip_rewards = 1
            d.clip_lo = _lib.NAN if self.clip.reward_min is None else float(self.clip.reward_min)
            d.clip_hi = _lib.NAN if self.clip.reward_max is None else float(self.clip.reward_max)
        if est is not None:
            value = st.view(sample_tree["agent_info"]["value"])
            boot = flat(st.view(sample_tree["bootstrap_value"]))
            adv = st.view(sample_tree["advantage"], writeback=True)
            ret = st.view(sample_tree["return_"], writeback=True)
            if value.dim() != 2:
                raise ValueError("the fused chain needs a [T, B] value estimate (no trailing agent axis)")
            keep.append(boot)
            d.value, d.ld_value, _ = rows_2d(value)
            d.bootstrap_value = boot.data_ptr()
            d.advantage, d.ld_advantage, _ = rows_2d(adv)
            d.return_, d.ld_return, _ = rows_2d(ret)
            d.estimate_advantage = 1
            if nr is not None and float(est.discount) != float(nr.discount):
                raise ValueError("NormalizeRewards and EstimateAdvantage must share the discount in the fused chain")
            d.discount = float(est.discount)
            d.gae_lambda = float(est.gae_lambda)
        elif na is not None:
            adv = ref if reward is None else st.view(sample_tree["advantage"], writeback=True)
            d.advantage, d.ld_advantage, _ = rows_2d(adv)
        if na is not None:
            d.normalize_advantage = 1
        d.eps_reward, d.eps_advantage = REWARD_EPSILON, ADV_EPSILON
        d.moments_flags = _lib.MOMENTS_F32_FLOW

        dev = st.device
        if self._ws is None or self._ws.device != dev:
            self._ws = torch.zeros(lib.plb_batch_pipeline_workspace_bytes(), dtype=torch.uint8, device=dev)
            self._moments = torch.zeros(6, dtype=torch.float64, device=dev)
        d.return_moments = self._moments.data_ptr()
        d.advantage_moments = self._moments.data_ptr() + 24
        d.workspace, d.workspace_bytes = self._ws.data_ptr(), self._ws.numel()

        # routes: on one GPU the library picks (PLAN_AUTO); on several, what ALL ranks can do
        distributed = self._distributed()
        in_kernel_exchange = False
        d.plan = _lib.PLAN_AUTO
        if distributed:
            d.plan = 0
            d.plan = self._agree_plan(lib, d, bytes(d) + bytes([phases]))
        # grid-exchange buffers: the resident kernel (one GPU when forced, or sharded) and the sharded two-kernel
        # route both trade their partial sums through them
        if phases == _lib.PHASE_ALL and (not st.staged or not distributed) and est is not None and na is not None \
                and (d.plan & (_lib.PLAN_RESIDENT | _lib.PLAN_CHUNKED)) and os.environ.get("PLB_EXCHANGE", "p2p") == "p2p":
            gx = self._grid_exchange(dev)
            if gx.ok:
                d.gx_peers, d.gx_rank, d.gx_world, d.gx_slots = gx.peer_table.data_ptr(), gx.rank, gx.world, gx.slots
                d.gx_self = gx.self_ptr
                keep.append(gx)
        if distributed and phases == _lib.PHASE_ALL and not st.staged and (d.plan & _lib.PLAN_CHUNKED):
            # the round-1 record-per-rank buffers: NormalizeRewards' statistics, and the fallback route
            from ..distributed import get_peer_exchange
            ex = get_peer_exchange(3, self.group, dev)
            if ex is not None and ex.world <= 32:
                d.xch_peers, d.xch_rank, d.xch_world = ex.peer_table.data_ptr(), ex.rank, ex.world
                keep.append(ex)
                in_kernel_exchange = True

        split_phases = distributed and phases == _lib.PHASE_ALL and not in_kernel_exchange
        carry_copies = nr is not None and T > 0 and not (d.carry_return_out and d.carry_done_out) \
            and (nr._carry_return is not None or nr._carry_done is not None)

        def enqueue():
            s = stream_ptr(dev)
            if split_phases:
                _lib.check(lib.plb_batch_pipeline_f32(ctypes.byref(d), _lib.PHASE_A, s), "fused transforms (A)")
                if nr is not None:
                    self._exchange(self._moments[0:3])
                _lib.check(lib.plb_batch_pipeline_f32(ctypes.byref(d), _lib.PHASE_B, s), "fused transforms (B)")
                if na is not None:
                    self._exchange(self._moments[3:6])
                _lib.check(lib.plb_batch_pipeline_f32(ctypes.byref(d), _lib.PHASE_C, s), "fused transforms (C)")
            else:
                _lib.check(lib.plb_batch_pipeline_f32(ctypes.byref(d), phases, s), "fused transforms")
            if carry_copies:
                if nr._carry_return is not None:
                    nr._carry_return.copy_(past[T - 1])
                if nr._carry_done is not None:
                    nr._carry_done.copy_(done[T - 1])

        # the NCCL all-gathers of the sharded chain are captured into the graph as well
        graphable = self.use_cuda_graph and not st.staged and not capturing and not any(snapshots)
        if self.launch == "direct" and self.use_cuda_graph and not st.staged and not split_phases and not carry_copies \
                and not any(snapshots):
            enqueue()
            self._bind_direct(sample_tree, phases, d, keep, dev)
        elif not graphable:
            enqueue()
        else:
            key = bytes(d) + bytes([phases])
            graph = self._graphs.get(key)
            if graph is not None:
                graph[0].replay()
                self._bind(sample_tree, phases, graph[0])
            elif key not in self._seen:
                self._seen.add(key)  # first sight: run eagerly (also configures the kernels)
                enqueue()
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    enqueue()
                g.replay()
                if len(self._graphs) > 64:
                    self._graphs.clear()
                self._graphs[key] = (g, keep, d)
        st.finish()
        return sample_tree

    def _leaf_checks(self, sample_tree):
        """(name, sub-name, leaf object, address) of every device leaf the chain may touch; None if one is on the host."""
        checks = []
        for name in self._LEAF_NAMES:
            if name in sample_tree:
                leaf = sample_tree[name]
                if self._leaf_signature(leaf) is None:
                    return None  # host leaf: never replayed
                checks.append((name, None, leaf, self._address(leaf)))
        if "agent_info" in sample_tree and "value" in sample_tree["agent_info"]:
            leaf = sample_tree["agent_info"]["value"]
            if self._leaf_signature(leaf) is None:
                return None
            checks.append(("agent_info", "value", leaf, self._address(leaf)))
        if len(self._bound) > 256:
            self._bound.clear()
        return checks

    def _bind_direct(self, sample_tree, phases: int, d, keep, dev) -> None:
        """Remember the filled descriptor for this tree object: while its leaves stay the same objects at the same
        addresses, a call is one C call (the descriptor holds raw device pointers; `keep` and the transforms'
        own buffers keep the auxiliary memory alive)."""
        checks = self._leaf_checks(sample_tree)
        if checks is None:
            return
        self._bound[(id(sample_tree), phases)] = (checks, None, (_lib.load().plb_batch_pipeline_f32, ctypes.byref(d), dev.index),
                                                  (d, keep))

    def _bind(self, sample_tree, phases: int, graph) -> None:
        """Remember the graph for this tree object together with the device leaves it was captured for."""
        checks = self._leaf_checks(sample_tree)
        if checks is None:
            return
        direct = (None, 0, 0)
        try:  # the instantiated graph's handle, when this torch exposes it
            exec_ptr = int(graph.raw_cuda_graph_exec())
            if exec_ptr and os.environ.get("PLB_DIRECT_GRAPH_LAUNCH", "1") == "1":
                direct = (_lib.load().plb_graph_launch, exec_ptr, torch.cuda.current_device())
        except Exception:
            pass
        self._bound[(id(sample_tree), phases)] = (checks, graph, direct)
