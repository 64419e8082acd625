"""Multi-GPU operation: the env axis ``B`` is sharded across the ranks of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The scans, the
elementwise passes and the replay gather shard with NO collective (independent env columns).
The only exchange step of the path is the statistics of the normalisers: each rank reduces its
shard to float64 ``(count, mean, M2)`` moments on the device, the moments are ALL-GATHERED
(24 bytes per rank for scalar statistics) and every rank merges them in rank order with the same
device kernel (``plb_merge_moments_f64``) -- an exact parallel (Chan) merge, bit-identical on
every rank, unlike the mean-of-variances all_reduce of parllel/torch/models.py:212-218.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._leaves import stream_ptr
from .transforms.norm_advantage import NormalizeAdvantage


def column_shard(B: int, world_size: int, rank: int) -> slice:
    """Contiguous block of env columns owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(B, world_size)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def all_gather_moments(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather a rank's float64 moments; returns ``[world, *local.shape]`` in rank order."""
    world = dist.get_world_size(group)
    flat = local.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat, group=group)  # flat in / flat out works for NCCL and gloo
    return out.reshape((world,) + tuple(local.shape))


def merge_gathered_moments(gathered: torch.Tensor, n_cols: int, layout: int, out: torch.Tensor) -> None:
    """Rank-ordered exact merge of gathered moments on the device (csrc/plb_stats.cu)."""
    rc = _lib.load().plb_merge_moments_f64(gathered.data_ptr(), gathered.shape[0], n_cols, layout, out.data_ptr(),
                                           stream_ptr(out.device))
    _lib.check(rc, "plb_merge_moments_f64")


class PeerExchange:
    """All-gather + rank-ordered merge of per-rank moments in ONE kernel over NVLink peer memory
    (csrc/plb_exchange.cu): direct P2P stores into every rank's symmetric buffer, release/acquire
    flags at system scope, no NCCL on the path.  The symmetric buffers come from
    ``torch.distributed._symmetric_memory`` (plumbing); the kernel is graph-capturable because its
    epoch counter lives in device memory.  One process per GPU."""

    def __init__(self, n_vals: int, group=None, device=None) -> None:
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_vals = int(n_vals)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        nbytes = _lib.load().plb_exchange_buffer_bytes(self.world, self.n_vals)
        self.buffer = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.buffer.zero_()
        self.handle = symm_mem.rendezvous(self.buffer, self.group)
        self.peer_table = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)  # every rank's buffer is zeroed and mapped before the first exchange

    def exchange(self, local: torch.Tensor, merged_out: torch.Tensor, layout: int, n_cols: int) -> None:
        assert local.numel() == self.n_vals and local.dtype == torch.float64 and local.is_contiguous()
        rc = _lib.load().plb_exchange_moments_p2p(local.data_ptr(), self.n_vals, self.peer_table.data_ptr(),
                                                  self.rank, self.world, layout, n_cols, merged_out.data_ptr(),
                                                  stream_ptr(self.device))
        _lib.check(rc, "plb_exchange_moments_p2p")


class GridExchange:
    """Buffers of the resident advantage kernel's in-kernel exchange (csrc/plb_scan_shared.cuh, GridXch):
    every CTA of every rank stores its partial moments into every rank's buffer and collects all of them
    from its own -- grid barrier and all-gather in one step.  With ``group`` None (one GPU) the buffer is
    (or ``local=True``: this GPU on its own) ordinary device memory; otherwise it is symmetric memory mapped into all peers (plumbing from
    ``torch.distributed._symmetric_memory``), and creating it is a collective.  ``ok`` is False when
    peer mapping is unavailable (the caller then keeps to the other routes)."""

    def __init__(self, group=None, device=None, local: bool = False) -> None:
        lib = _lib.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.slots = int(lib.plb_grid_exchange_slots())
        self.ok = self.slots > 0
        self.handle = None
        self.self_ptr = 0
        single = local or not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1)
        if single:
            self.group, self.world, self.rank = None, 1, 0
            self.buffer = torch.zeros(lib.plb_grid_exchange_bytes(1, max(self.slots, 1)), dtype=torch.uint8, device=self.device)
            self.peer_table = torch.tensor([self.buffer.data_ptr()], dtype=torch.int64, device=self.device)
            self.self_ptr = self.buffer.data_ptr()
            return
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = lib.plb_grid_exchange_bytes(self.world, max(self.slots, 1))
            if nbytes == 0:
                raise RuntimeError(f"unsupported world size {self.world}")
            self.buffer = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.buffer.zero_()
            self.handle = symm_mem.rendezvous(self.buffer, self.group)
            self.peer_table = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=self.device)
            self.self_ptr = int(self.handle.buffer_ptrs[self.rank])
            torch.cuda.synchronize(self.device)
            good = 1
        except Exception as err:  # no peer mapping on this system
            import warnings
            warnings.warn(f"grid exchange buffers unavailable ({err!r}); using the two-kernel route")
            good = 0
        # all ranks take the resident route or none does
        flag = torch.tensor([good], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)  # also: every buffer is zeroed and mapped
        self.ok = self.ok and bool(flag.item())

    def status(self) -> int:
        """0 = every launch so far saw all ranks; 1 = one timed out waiting (its results are NaN)."""
        return int(_lib.load().plb_grid_exchange_status(self.buffer.data_ptr(), stream_ptr(self.device)))


_peer_exchanges: dict = {}
_peer_exchange_failed = False


def get_peer_exchange(n_vals: int, group, device) -> "PeerExchange | None":
    """The cached peer-memory exchange for (group, n_vals), or None when symmetric memory is not
    available / ``PLB_EXCHANGE=nccl`` is set."""
    import os
    global _peer_exchange_failed
    if os.environ.get("PLB_EXCHANGE", "p2p") != "p2p" or _peer_exchange_failed:
        return None
    key = (id(group), int(n_vals), torch.device(device).index)
    ex = _peer_exchanges.get(key)
    if ex is None:
        try:
            ex = PeerExchange(n_vals, group, device)
            _peer_exchanges[key] = ex
        except Exception as err:  # symmetric memory unavailable on this system: fall back to NCCL
            import warnings
            warnings.warn(f"peer-memory exchange unavailable ({err!r}); using NCCL all-gather")
            _peer_exchange_failed = True
            return None
    return ex


def exchange_moments(moments: torch.Tensor, n_cols: int, layout: int, group=None) -> None:
    """In-place exchange of a rank's moments: afterwards ``moments`` holds the rank-ordered merge over
    the whole group.  Uses the peer-memory kernel when symmetric memory is available (default; set
    ``PLB_EXCHANGE=nccl`` to force the NCCL all-gather + merge-kernel path), else NCCL."""
    import os
    global _peer_exchange_failed
    use_p2p = os.environ.get("PLB_EXCHANGE", "p2p") == "p2p" and moments.is_cuda and not _peer_exchange_failed \
        and moments.numel() <= 4096
    if use_p2p:
        key = (id(group), moments.numel(), moments.device.index)
        ex = _peer_exchanges.get(key)
        if ex is None:
            try:
                ex = PeerExchange(moments.numel(), group, moments.device)
                _peer_exchanges[key] = ex
            except Exception as err:  # symmetric memory unavailable on this system: fall back to NCCL
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({err!r}); using NCCL all-gather")
                _peer_exchange_failed = True
                ex = None
        if ex is not None:
            ex.exchange(moments, moments, layout, n_cols)
            return
    gathered = all_gather_moments(moments, group)
    merge_gathered_moments(gathered, n_cols, layout, moments)


class ShardedNormalizeAdvantage(NormalizeAdvantage):
    """``NormalizeAdvantage`` over a B-sharded batch: local moments -> all-gather -> merge ->
    local normalise.  Every rank normalises with the statistics of the WHOLE batch."""

    def __init__(self, sample_tree, group=None) -> None:
        super().__init__(sample_tree)
        self.group = group

    def _exchange(self, moments: torch.Tensor, inner: int) -> None:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        exchange_moments(moments[: 3 * inner], inner, 0, self.group)


def shard_statistics(transform, group=None):
    """Switch a statistics-carrying transform (``NormalizeRewards``, ``NormalizeObservations``) to
    B-sharded operation: every update all-gathers the per-rank batch moments over ``group`` and merges
    them in rank order, so all ranks hold bit-identical running statistics of the WHOLE batch.
    Returns the transform for chaining."""
    for name in ("return_statistics", "obs_statistics"):
        stats = getattr(transform, name, None)
        if stats is not None:
            stats.sharded = True
            stats.group = group
    return transform
