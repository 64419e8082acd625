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
        """Stores more than a batch's worth of samples in a circular buffer for off-policy
        algorithms to sample from.

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
        self._cursor: int = 0  # index of next sample to write
        self._full = False  # has the entire buffer been written to at least once?
        if batch_transform is None:
            batch_transform = lambda x: x  # noqa: E731
        self.batch_transform = batch_transform
        self.seed()

        self._paths, self._host, self._ring = [], [], []
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
            if self.device is None:
                self.device = ring.device
            assert ring.shape[0] == size_T and ring.shape[1] == sampler_batch_spec.B, \
                f"leaf {'.'.join(path)}: expected leading dims ({size_T}, {sampler_batch_spec.B}), got {tuple(ring.shape[:2])}"
            for size, stride, want in zip(reversed(ring.shape[2:]), reversed(ring.stride()[2:]),
                                          np.cumprod((1,) + tuple(reversed(ring.shape[2:])))):
                assert size == 1 or stride == want, "feature axes of a replay leaf must be contiguous"
            self._host.append(host)
            self._ring.append(ring)
        if len(self._ring) > _lib.GATHER_MAX_LEAVES:
            raise ValueError(f"at most {_lib.GATHER_MAX_LEAVES} leaves per gather launch")
        # pinned staging for the sampled indices (T then B), reused by every call
        self._idx_host = None
        self._idx_dev = None

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
        """The reference's index math (replay.py:55-70): int64 ``(T_idxs, B_idxs)`` on the host."""
        n = self.batch_size if batch_size is None else batch_size
        if self._full:
            # valid region for sampling wraps around: sample [0, L) then offset and wrap
            offset = self._cursor + self.oldest_n_samples_invalid
            valid_length = self.size_T - self.oldest_n_samples_invalid - self.newest_n_samples_invalid
            T_idxs = self._rng.integers(0, valid_length, size=(n,))
            T_idxs = (T_idxs + offset) % self.size_T
        else:
            valid_length = self._cursor - self.newest_n_samples_invalid
            T_idxs = self._rng.integers(0, valid_length, size=(n,))
        B_idxs = self._rng.integers(0, self.batch_spec.B, size=(n,))
        return T_idxs, B_idxs

    def gather(self, T_idxs, B_idxs) -> ArrayDict:
        """``self.tree[T_idxs, B_idxs]`` (replay.py:72) as one CUDA launch; indices may be host
        int64 arrays or CUDA int64 tensors.  Returns freshly allocated ``[n, *feat]`` tensors."""
        dev = self.device
        if isinstance(T_idxs, torch.Tensor) and T_idxs.is_cuda:
            t_dev, b_dev = T_idxs.contiguous(), B_idxs.contiguous()
            n = t_dev.numel()
        else:
            n = len(T_idxs)
            if self._idx_host is None or self._idx_host.shape[1] < n:
                self._idx_host = torch.empty((2, n), dtype=torch.int64).pin_memory()
                self._idx_dev = torch.empty((2, n), dtype=torch.int64, device=dev)
            self._idx_host[0, :n] = torch.from_numpy(np.ascontiguousarray(T_idxs, np.int64))
            self._idx_host[1, :n] = torch.from_numpy(np.ascontiguousarray(B_idxs, np.int64))
            self._idx_dev[:, :n].copy_(self._idx_host[:, :n], non_blocking=True)
            t_dev, b_dev = self._idx_dev[0, :n], self._idx_dev[1, :n]
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
        # rows the sampler has just written: [cursor, cursor + T) -- refresh the HBM mirror
        lo, hi = self._cursor, self._cursor + self.batch_spec.T
        for host, ring in zip(self._host, self._ring):
            if host is not None:
                ring[lo:hi].copy_(host[lo:hi], non_blocking=False)
        # move cursor forward
        self._cursor += self.batch_spec.T
        if self._cursor >= self.size_T:
            self._cursor %= self.size_T
            self._full = True
