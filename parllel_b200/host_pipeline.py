"""Host-side execution of the batch transforms: pinned leaves and a software-pipelined H2D /
compute / D2H path for trees that live in host memory (the reference's ``ArrayDict[Array]``).

``PinnedBatch`` allocates a sample tree on page-locked memory (the same allocation
parllel/patterns.py:104-235 makes, DMA-friendly).  ``HostBatchPipeline`` wraps a
``FusedBatchTransforms`` and runs it on host trees through two device-resident slots and three
CUDA streams, so that the upload of batch ``i+1``, the kernels of batch ``i`` and the download of
batch ``i-1`` overlap (PCIe is full duplex); results land in the caller's own host buffers.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from .tree import ArrayDict

_registered: dict[int, int] = {}  # host pointer -> nbytes registered with cudaHostRegister


def pinned_like(array: np.ndarray) -> np.ndarray:
    """A page-locked numpy array with the contents of ``array`` (owns its pinned torch storage)."""
    t = torch.empty(array.shape, dtype=torch.from_numpy(np.empty(0, array.dtype)).dtype).pin_memory()
    out = t.numpy()
    out[...] = array
    _PinnedHolder(out, t)
    return out


def pin_in_place(array: np.ndarray) -> bool:
    """Page-lock an existing host buffer (``cudaHostRegister``) so copies to/from it are direct DMA.
    Returns False (and leaves the buffer pageable) if registration is refused."""
    base = array
    while isinstance(base.base, np.ndarray):
        base = base.base
    ptr, nbytes = base.ctypes.data, base.nbytes
    if torch.from_numpy(base.reshape(-1)[:1]).is_pinned() or _registered.get(ptr, 0) >= nbytes:
        return True
    rc = torch.cuda.cudart().cudaHostRegister(ptr, nbytes, 0)
    if int(rc) != 0:
        return False
    _registered[ptr] = nbytes
    # the registration must not outlive the buffer: unregister when numpy frees it
    weakref.finalize(base, _unregister, ptr)
    return True


def _unregister(ptr: int) -> None:
    if _registered.pop(ptr, None) is not None:
        try:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(ptr)
        except Exception:
            pass


class _PinnedHolder:
    """Keeps the owning pinned tensor alive for as long as the process lives."""
    _alive: list = []

    def __init__(self, array: np.ndarray, tensor: torch.Tensor) -> None:
        self.array, self.tensor = array, tensor
        _PinnedHolder._alive.append(self)


class PinnedBatch:
    """A [T, B] sample tree (reward, done, agent_info.value, bootstrap_value, advantage, return_)
    on page-locked host memory, filled from a dict of numpy arrays."""

    def __init__(self, batch: dict) -> None:
        value = batch["value"]
        self.tree = ArrayDict({
            "reward": pinned_like(batch["reward"]),
            "done": pinned_like(batch["done"]),
            "agent_info": {"value": pinned_like(value)},
            "bootstrap_value": pinned_like(batch["bootstrap_value"]),
            "advantage": pinned_like(np.zeros_like(value)),
            "return_": pinned_like(np.zeros_like(value)),
        })


class _Handle:
    def __init__(self, events, tree) -> None:
        self._events, self.tree = events, tree

    def wait(self):
        """Block until this batch's results are in the caller's host buffers; returns the tree."""
        for ev in self._events:
            ev.synchronize()
        return self.tree


_INPUTS = ("reward", "done", "bootstrap_value", "valid")
_OUTPUTS_ALWAYS = ("advantage", "return_")


class HostBatchPipeline:
    """``pipe.submit(tree)`` enqueues upload -> fused transforms -> download for a host tree and
    returns a handle; ``handle.wait()`` completes it.  ``pipe(tree)`` is submit + wait (drop-in
    ``Transform`` behaviour).  Submit batch ``i+1`` before waiting on batch ``i`` to overlap the
    transfers of consecutive batches.  Leaves must be numpy-viewable; they are page-locked in place
    on first use.  The carry rows of ``NormalizeRewards`` are kept by the transform itself (the
    batches are consecutive by construction)."""

    def __init__(self, fused, device=None, n_slots: int = 2) -> None:
        from ._leaves import default_device
        self.fused = fused
        self.device = torch.device(device) if device is not None else default_device()
        self.n_slots = n_slots
        self._slots: list = []
        self._views: dict = {}
        self._slot_free: list = []
        self._next = 0
        # two streams per direction: consecutive copies on ONE stream leave the link idle between transfers (each
        # waits for its predecessor's completion before its DMA is set up: ~25 us per leaf, 10 % of a C1 step);
        # leaves alternate between the two, so one transfer's set-up hides behind the other's payload
        self._h2d = [torch.cuda.Stream(self.device) for _ in range(2)]
        self._compute = torch.cuda.Stream(self.device)
        self._d2h = [torch.cuda.Stream(self.device) for _ in range(2)]
        self._outputs = list(_OUTPUTS_ALWAYS)
        if fused.norm_rewards is not None:
            self._outputs += ["reward", "past_return"]
        elif fused.clip is not None:
            self._outputs += ["reward"]

    def _host_view(self, leaf) -> torch.Tensor:
        """Torch view of a host leaf, page-locked in place; remembered per leaf object (the samplers hand the same
        buffers over every iteration), so a submit costs no pointer-attribute queries."""
        hit = self._views.get(id(leaf))
        if hit is not None and hit[0] is leaf:
            return hit[1]
        arr = leaf.detach().numpy() if isinstance(leaf, torch.Tensor) else np.asarray(leaf)
        pin_in_place(arr)
        view = torch.from_numpy(arr)
        if len(self._views) > 256:
            self._views.clear()
        self._views[id(leaf)] = (leaf, view)  # holding the leaf keeps its id from being reused
        return view

    def _make_slot(self, tree) -> ArrayDict:
        def dev_like(leaf):
            h = self._host_view(leaf)
            return torch.empty(h.shape, dtype=h.dtype, device=self.device)
        names = [n for n in ("reward", "done", "bootstrap_value", "valid", "advantage", "return_", "past_return")
                 if n in tree]
        slot = ArrayDict({n: dev_like(tree[n]) for n in names})
        slot["agent_info"] = ArrayDict({"value": dev_like(tree["agent_info"]["value"])})
        return slot

    def submit(self, tree) -> _Handle:
        k = self._next % self.n_slots
        self._next += 1
        if len(self._slots) <= k:
            self._slots.append(self._make_slot(tree))
            self._slot_free.append(None)
        slot = self._slots[k]
        uploads = [(slot[n], self._host_view(tree[n])) for n in _INPUTS if n in tree]
        uploads.append((slot["agent_info"]["value"], self._host_view(tree["agent_info"]["value"])))
        uploads.sort(key=lambda pair: -pair[1].numel() * pair[1].element_size())  # big leaves first, one per stream
        uploaded = []
        for j, stream in enumerate(self._h2d):
            with torch.cuda.stream(stream):
                if self._slot_free[k] is not None:
                    for ev in self._slot_free[k]:
                        stream.wait_event(ev)  # previous user of the slot has been downloaded
                for dst, src in uploads[j::len(self._h2d)]:
                    dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
                uploaded.append(ev)
        with torch.cuda.stream(self._compute):
            for ev in uploaded:
                self._compute.wait_event(ev)
            self.fused(slot)
            computed = torch.cuda.Event()
            computed.record(self._compute)
        downloads = [(self._host_view(tree[n]), slot[n]) for n in self._outputs if n in tree]
        done = []
        for j, stream in enumerate(self._d2h):
            with torch.cuda.stream(stream):
                stream.wait_event(computed)
                for dst, src in downloads[j::len(self._d2h)]:
                    dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
                done.append(ev)
        self._slot_free[k] = done
        return _Handle(done, tree)

    def __call__(self, tree):
        return self.submit(tree).wait()
