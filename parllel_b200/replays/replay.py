"""``ReplayBuffer`` on the B200 (parllel/replays/replay.py:12-93).

The ring lives in HBM and ``sample_batch()`` is ONE gather launch over every leaf
(csrc/plb_gather.cu).  Index generation follows the reference bit for bit: the same
``numpy.random.Generator`` (PCG64) calls in the same order -- T draw, then B draw.

Ring residency:
  * leaves that are CUDA tensors / ``DeviceArray`` are used in place (zero copy): whatever
    writes the samples (a device-side sampler) is immediately visible, as in the reference
    where the replay tree shares memory with the sampler's arrays;
  * host leaves (numpy arrays, CPU tensors sharing memory with parllel ``Array``s, as
    examples/continuous_cartpole/train_sac.py:122-123 builds them) are mirrored in HBM once and
    ``next_iteration()`` uploads only the ``T`` rows the sampler wrote since the last call.
"""
from __future__ import annotations

from typing import Callable, Iterator, Optional

import numpy as np
import torch
from numpy import random

from .. import _lib
from .._leaves import default_device, stream_ptr
from ..arrays import DeviceArray
from ..tree import ArrayDict
from ..spec import BatchSpec


def _flatten(tree, prefix=()):
    for key in tree:
        leaf = tree[key]
        if hasattr(leaf, "keys") and not hasattr(leaf, "shape"):
            yield from _flatten(leaf, prefix + (key,))
        else:
            yield prefix + (key,), leaf


def _unflatten(pairs) -> ArrayDict:
    root: dict = {}
    for path, leaf in pairs:
        node = root
        for key in path[:-1]:
            node = node.setdefault(key, {})
        node[path[-1]] = leaf
    return ArrayDict(root)


class ReplayBuffer:
    def __init__(
        self,
        tree,
        sampler_batch_spec: BatchSpec,
        size_T: int,
        replay_batch_size: int,
        newest_n_samples_invalid: int = 0,
        oldest_n_samples_invalid: int = 0,
        batch_transform: Optional[Callable] = None,
        device=None,
        device_rng: bool = False,
    ) -> None:
        """Ring of ``size_T`` time steps x ``B`` environments in HBM that off-policy algorithms draw
        uniform minibatches from (interface of parllel/replays/replay.py:12-52).

        ``device_rng=True`` draws the indices on the GPU (csrc/plb_rng.cu, bit-exact with the host
        ``numpy.random.Generator`` stream) so a minibatch needs no host->device copy at all."""
        self.device_rng = device_rng
        self._rng_state = None
        self.tree = tree
        self.batch_spec = sampler_batch_spec
        self.size_T = size_T
        self.batch_size = replay_batch_size
        self.newest_n_samples_invalid = newest_n_samples_invalid
        self.oldest_n_samples_invalid = oldest_n_samples_invalid
        # ring position: `_cursor` is the row the sampler's next batch lands on, `_full` turns true at the
        # first wrap-around (same two attributes, same meaning, as replay.py:41-42 -- tests poke them)
        self._cursor: int = 0
        self._full = False
        self.batch_transform = batch_transform if batch_transform is not None else (lambda batch: batch)
        self.seed()

        self._paths, self._host, self._ring, self._host_pinned = [], [], [], []
        self.device = torch.device(device) if device is not None else None
        for path, leaf in _flatten(tree):
            self._paths.append(path)
            if isinstance(leaf, DeviceArray):
                ring, host = leaf.tensor, None
            elif isinstance(leaf, torch.Tensor) and leaf.is_cuda:
                ring, host = leaf, None
            else:
                host = leaf if isinstance(leaf, torch.Tensor) else torch.from_numpy(np.asarray(leaf))
                if self.device is None:
                    self.device = default_device()
                ring = host.to(self.device).contiguous()  # one-time mirror of the whole ring
                # page-lock the caller's ring in place so the per-iteration refresh is an asynchronous DMA
                # instead of a blocking pageable copy per leaf (falls back silently if registration is refused)
                pinned = False
                if ring.is_cuda:
                    from ..host_pipeline import pin_in_place
                    pinned = bool(host.is_pinned() or pin_in_place(host.numpy()))
                self._host_pinned.append(pinned)
            if self.device is None:
                self.device = ring.device
            assert ring.shape[0] == size_T and ring.shape[1] == sampler_batch_spec.B, \
                f"leaf {'.'.join(path)}: expected leading dims ({size_T}, {sampler_batch_spec.B}), got {tuple(ring.shape[:2])}"
            for size, stride, want in zip(reversed(ring.shape[2:]), reversed(ring.stride()[2:]),
                                          np.cumprod((1,) + tuple(reversed(ring.shape[2:])))):
                assert size == 1 or stride == want, "feature axes of a replay leaf must be contiguous"
            self._host.append(host)
            self._ring.append(ring)
            if host is None:
                self._host_pinned.append(False)
        if len(self._ring) > _lib.GATHER_MAX_LEAVES:
            raise ValueError(f"at most {_lib.GATHER_MAX_LEAVES} leaves per gather launch")
        # host-drawn indices travel through a small ring of pinned staging buffers; each slot carries the
        # event of the upload that last used it, so a backlogged stream can never see a later call's indices
        self._idx_slots: list = []
        self._idx_next = 0

    def seed(self, seed: Optional[int] = None) -> None:
        self._rng = random.default_rng(seed)
        self._rng_state = None  # device copy of the PCG64 state is rebuilt lazily

    def _device_rng_state(self) -> torch.Tensor:
        """PCG64 state words in device memory, seeded from the numpy generator (SeedSequence hashing
        stays numpy's; only the stream is continued on the device)."""
        if self._rng_state is None:
            st = self._rng.bit_generator.state
            assert st["bit_generator"] == "PCG64"
            s, inc, m = st["state"]["state"], st["state"]["inc"], (1 << 64) - 1
            words = [s >> 64, s & m, inc >> 64, inc & m, (st["has_uint32"] & 1) | (int(st["uinteger"]) << 32)]
            host = np.array(words, dtype=np.uint64).view(np.int64)
            self._rng_state = torch.from_numpy(host.copy()).to(self.device)
        return self._rng_state

    def sample_indices_device(self, k: int = 1):
        """``k`` minibatches of ``(T_idxs, B_idxs)`` drawn on the device: int64 CUDA tensors ``[k, n]``."""
        n = self.batch_size
        t = torch.empty((k, n), dtype=torch.int64, device=self.device)
        b = torch.empty((k, n), dtype=torch.int64, device=self.device)
        rc = _lib.load().plb_replay_sample_indices(
            self._device_rng_state().data_ptr(), self.size_T, self.batch_spec.B, self._cursor, int(self._full),
            self.newest_n_samples_invalid, self.oldest_n_samples_invalid, n, k, t.data_ptr(), b.data_ptr(),
            stream_ptr(self.device))
        _lib.check(rc, "plb_replay_sample_indices")
        return t, b

    @property
    def replay_batch_size(self) -> int:
        return self.batch_size

    @property
    def capacity(self):
        return self.size_T * self.batch_spec.B

    # ------------------------------------------------------------------------------------
    def sample_indices(self, batch_size: Optional[int] = None):
        """Host-side draw with the reference's generator calls in the reference's order (replay.py:55-70):
        one ``integers`` call for the time indices, then one for the environment indices; int64 arrays."""
        n = self.batch_size if batch_size is None else batch_size
        invalid = self.oldest_n_samples_invalid + self.newest_n_samples_invalid
        if self._full:
            # every row is populated; the usable window starts `oldest_invalid` rows after the write position
            # and is addressed modulo the ring length
            first_valid = self._cursor + self.oldest_n_samples_invalid
            T_idxs = (self._rng.integers(0, self.size_T - invalid, size=(n,)) + first_valid) % self.size_T
        else:
            # only rows [0, cursor) exist yet, minus the newest ones still being completed
            T_idxs = self._rng.integers(0, self._cursor - self.newest_n_samples_invalid, size=(n,))
        B_idxs = self._rng.integers(0, self.batch_spec.B, size=(n,))
        return T_idxs, B_idxs

    _IDX_SLOTS = 4

    def _stage_indices(self, T_idxs, B_idxs, n: int):
        """Upload host indices through the next pinned slot; returns the device views ``(t, b)``."""
        dev = self.device
        k = self._idx_next % self._IDX_SLOTS
        self._idx_next += 1
        if len(self._idx_slots) <= k:
            self._idx_slots.append(None)
        slot = self._idx_slots[k]
        if slot is None or slot[0].shape[1] < n:
            slot = [torch.empty((2, n), dtype=torch.int64).pin_memory(),
                    torch.empty((2, n), dtype=torch.int64, device=dev), None]
            self._idx_slots[k] = slot
        host, devbuf, uploaded = slot
        if uploaded is not None:
            uploaded.synchronize()  # the DMA that last read this pinned slot has finished
        host[0, :n] = torch.from_numpy(np.ascontiguousarray(T_idxs, np.int64))
        host[1, :n] = torch.from_numpy(np.ascontiguousarray(B_idxs, np.int64))
        devbuf[:, :n].copy_(host[:, :n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        slot[2] = ev
        return devbuf[0, :n], devbuf[1, :n]

    def gather(self, T_idxs, B_idxs) -> ArrayDict:
        """``self.tree[T_idxs, B_idxs]`` (replay.py:72) as one CUDA launch; indices may be host
        int64 arrays or CUDA int64 tensors.  Returns freshly allocated ``[n, *feat]`` tensors."""
        dev = self.device
        if isinstance(T_idxs, torch.Tensor) and T_idxs.is_cuda:
            t_dev, b_dev = T_idxs.contiguous(), B_idxs.contiguous()
            n = t_dev.numel()
        else:
            n = len(T_idxs)
            t_dev, b_dev = self._stage_indices(T_idxs, B_idxs, n)
        outs = []
        leaves = (_lib.GatherLeaf * len(self._ring))()
        for k, ring in enumerate(self._ring):
            out = torch.empty((n,) + tuple(ring.shape[2:]), dtype=ring.dtype, device=dev)
            outs.append(out)
            item = ring.element_size()
            row_bytes = int(np.prod(ring.shape[2:], dtype=np.int64)) * item
            leaves[k] = _lib.GatherLeaf(ring.data_ptr(), out.data_ptr(), row_bytes,
                                        ring.stride(0) * item, ring.stride(1) * item)
        rc = _lib.load().plb_gather_tree(leaves, len(self._ring), t_dev.data_ptr(), b_dev.data_ptr(), n,
                                         stream_ptr(dev))
        _lib.check(rc, "plb_gather_tree")
        return _unflatten(zip(self._paths, outs))

    def sample_batch(self) -> ArrayDict:
        if self.device_rng:
            t, b = self.sample_indices_device(1)
            T_idxs, B_idxs = t[0], b[0]
        else:
            T_idxs, B_idxs = self.sample_indices()
        batch = self.gather(T_idxs, B_idxs)
        batch = self.batch_transform(batch)
        return batch

    def sample_batches(self, k: int) -> ArrayDict:
        """``k`` consecutive ``sample_batch()`` calls fused into ONE gather launch.  The index
        stream is exactly that of ``k`` sequential calls (T draw then B draw per minibatch), which is
        legal whenever the ring is not written in between -- SAC draws ``updates_per_optimize``
        minibatches per iteration that way (torch/algos/sac.py:112-115).  Leaves come back as
        ``[k, batch, *feat]``; ``batch_transform`` is applied to the stacked tree."""
        if self.device_rng:
            t, b = self.sample_indices_device(k)
            flat = self.gather(t.reshape(-1), b.reshape(-1))
        else:
            idx = [self.sample_indices() for _ in range(k)]
            T_all = np.concatenate([t for t, _ in idx])
            B_all = np.concatenate([b for _, b in idx])
            flat = self.gather(T_all, B_all)
        n = self.batch_size
        stacked = flat.apply(lambda leaf: leaf.reshape((k, n) + tuple(leaf.shape[1:])))
        return self.batch_transform(stacked)

    def batches(self) -> Iterator[ArrayDict]:
        while True:
            yield self.sample_batch()

    def __iter__(self) -> Iterator[ArrayDict]:
        # the reference's `yield from self.batches` (replay.py:80-81) omits the call and raises
        # TypeError when iterated; the evident intent is implemented here.
        yield from self.batches()

    def next_iteration(self) -> None:
        """The sampler has filled rows [cursor, cursor + T): bring the HBM mirror of host leaves up to date
        (device leaves are shared with the writer and need nothing), then advance the write position."""
        T = self.batch_spec.T
        lo = self._cursor
        for host, ring, pinned in zip(self._host, self._ring, self._host_pinned):
            if host is not None:
                # the sampler will not touch these rows again before a whole revolution of the ring, so a
                # stream-ordered DMA out of pinned memory needs no host-side wait
                ring[lo:lo + T].copy_(host[lo:lo + T], non_blocking=pinned)
        self._cursor = lo + T
        if self._cursor >= self.size_T:  # wrapped: from now on the whole ring holds samples
            self._cursor -= self.size_T * (self._cursor // self.size_T)
            self._full = True
