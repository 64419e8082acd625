"""``BatchedDataLoader`` on the B200 (parllel/replays/batched_dataloader.py:14-121).

PPO's consumer of the transformed batch: every epoch shuffles ``arange(T*B)`` (or ``arange(B)`` for
recurrent agents), splits it into ``n_batches`` pieces and gathers each piece out of the sample
tree.  The shuffle is the reference's own ``numpy.random.Generator.shuffle`` call on the host
(bit-exact index order); every minibatch is ONE launch of the replay gather kernel
(csrc/plb_gather.cu) over all leaves of a device-resident tree, so the whole-tree H2D copy the
reference makes per epoch (``pre_batches_transform = tree.to(device)``, train_ppo.py:142) happens at
most once per epoch here (host trees) or never (device trees).
"""
from __future__ import annotations

from typing import Callable, Iterator

import numpy as np
import torch

from .. import _lib
from .._leaves import default_device, stream_ptr
from ..arrays import DeviceArray
from ..tree import ArrayDict
from ..spec import BatchSpec
from .replay import _flatten, _unflatten


class BatchedDataLoader:
    """Iterates through a tree of samples in a fixed number of batches.  Fields listed in
    ``batch_only_fields`` (e.g. ``initial_rnn_state``) have no time axis and are indexed by env only."""

    def __init__(
        self,
        tree,
        sampler_batch_spec: BatchSpec,
        n_batches: int,
        batch_only_fields: list[str] | None = None,
        recurrent: bool = False,
        shuffle: bool = True,
        drop_last: bool = True,
        pre_batches_transform: Callable | None = None,
        batch_transform: Callable | None = None,
        device=None,
    ) -> None:
        identity = lambda tree_: tree_  # noqa: E731
        self.source_tree = tree
        self.tree = tree
        self.sampler_batch_spec = sampler_batch_spec
        self.batch_only_fields = batch_only_fields
        self.recurrent, self.shuffle, self.drop_last = recurrent, shuffle, drop_last
        self.n_batches = n_batches
        # unit of shuffling: one environment's whole trajectory for recurrent agents, one (t, b) sample otherwise
        self.size = sampler_batch_spec.B if recurrent else sampler_batch_spec.T * sampler_batch_spec.B
        self.batch_size = self.size // n_batches
        self.pre_batches_transform = pre_batches_transform or identity
        self.batch_transform = batch_transform or identity
        self.device = torch.device(device) if device is not None else None
        self.seed(None)

    def seed(self, seed: int | None) -> None:
        self.rng: np.random.Generator = np.random.default_rng(seed)

    def __len__(self) -> int:
        return self.size

    # ------------------------------------------------------------------------------------
    def _device_leaves(self, tree):
        """(paths, [T, B, ...] or [B, ...] CUDA tensors, is_batch_only flags) of the current tree."""
        paths, rings, batch_only = [], [], []
        only = set(self.batch_only_fields or ())
        for path, leaf in _flatten(tree):
            if isinstance(leaf, DeviceArray):
                t = leaf.tensor
            elif isinstance(leaf, torch.Tensor):
                t = leaf
            else:
                t = torch.from_numpy(np.asarray(leaf))
            if not t.is_cuda:
                if self.device is None:
                    self.device = default_device()
                t = t.to(self.device, non_blocking=True)
            if self.device is None:
                self.device = t.device
            paths.append(path)
            rings.append(t)
            batch_only.append(path[0] in only)
        return paths, rings, batch_only

    def _gather(self, paths, rings, batch_only, t_idx: torch.Tensor, b_idx: torch.Tensor, lead_shape) -> ArrayDict:
        n = t_idx.numel()
        leaves = (_lib.GatherLeaf * len(rings))()
        outs = []
        for k, (ring, only_b) in enumerate(zip(rings, batch_only)):
            feat = tuple(ring.shape[1:] if only_b else ring.shape[2:])
            item = ring.element_size()
            row_bytes = max(int(np.prod(feat, dtype=np.int64)), 1) * item
            shape = ((lead_shape[-1],) if only_b else tuple(lead_shape)) + feat
            out = torch.empty(shape, dtype=ring.dtype, device=ring.device)
            outs.append(out)
            if only_b:  # no time axis: every (t, b) pair maps to row b; gather the first lead_shape[-1] samples
                leaves[k] = _lib.GatherLeaf(ring.data_ptr(), out.data_ptr(), row_bytes, 0, ring.stride(0) * item)
            else:
                leaves[k] = _lib.GatherLeaf(ring.data_ptr(), out.data_ptr(), row_bytes,
                                            ring.stride(0) * item, ring.stride(1) * item)
        lib = _lib.load()
        s = stream_ptr(self.device)
        time_major = [k for k, only_b in enumerate(batch_only) if not only_b]
        env_only = [k for k, only_b in enumerate(batch_only) if only_b]
        if time_major:
            sub = (_lib.GatherLeaf * len(time_major))(*[leaves[k] for k in time_major])
            _lib.check(lib.plb_gather_tree(sub, len(time_major), t_idx.data_ptr(), b_idx.data_ptr(), n, s),
                       "plb_gather_tree")
        if env_only:
            nb = lead_shape[-1]
            sub = (_lib.GatherLeaf * len(env_only))(*[leaves[k] for k in env_only])
            # the last `nb` entries of b_idx enumerate the batch's envs once (see batches())
            _lib.check(lib.plb_gather_tree(sub, len(env_only), t_idx[n - nb:].data_ptr(), b_idx[n - nb:].data_ptr(), nb, s),
                       "plb_gather_tree")
        return _unflatten(zip(paths, outs))

    def _pieces(self):
        """(start, stop) of every minibatch in the permuted index list: ``n_batches`` equal pieces, plus a
        shorter final one when the size does not divide and ``drop_last`` is off."""
        stop_at = self.size - self.batch_size + 1 if self.drop_last else self.size
        return [(lo, min(lo + self.batch_size, self.size)) for lo in range(0, stop_at, self.batch_size)]

    def batches(self) -> Iterator[ArrayDict]:
        self.tree = self.pre_batches_transform(self.source_tree)
        paths, rings, batch_only = self._device_leaves(self.tree)
        T = self.sampler_batch_spec.T

        # the permutation is drawn exactly as the reference draws it (int32 arange, Generator.shuffle), so a
        # seeded loader visits the samples in the reference's order; it is uploaded once per epoch
        order = np.arange(self.size, dtype=np.int32)
        if self.shuffle:
            self.rng.shuffle(order)
        perm = torch.from_numpy(order.astype(np.int64)).to(self.device, non_blocking=True)
        time_steps = torch.arange(T, device=self.device, dtype=torch.int64) if self.recurrent else None
        for lo, hi in self._pieces():
            picked = perm[lo:hi]
            nb = hi - lo
            if self.recurrent:
                # whole trajectories of the picked environments: out[t, i] = leaf[t, picked[i]]
                t_idx = time_steps.repeat_interleave(nb)
                b_idx = picked.repeat(T)
                batch = self._gather(paths, rings, batch_only, t_idx, b_idx, (T, nb))
            else:
                # a flat sample number s addresses (t, b) = (s mod T, s div T); leading dims collapse to [nb]
                t_idx = torch.remainder(picked, T)
                b_idx = torch.div(picked, T, rounding_mode="floor")
                batch = self._gather(paths, rings, batch_only, t_idx, b_idx, (nb,))
            yield self.batch_transform(batch)

    def __iter__(self) -> Iterator[ArrayDict]:
        yield from self.batches()
