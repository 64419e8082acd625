"""Leaf resolution at the drop-in boundary.

A transform receives a tree whose leaves are one of
  * ``torch.Tensor`` on a CUDA device           -> used in place, zero copy (fast path);
  * :class:`parllel_b200.arrays.DeviceArray`    -> its window view, zero copy (fast path);
  * anything numpy can view (parllel ``Array``, ``np.ndarray``, CPU tensor)
                                                -> staged: H2D copy, kernel, D2H write-back
                                                   into the caller's buffer (drop-in path).
The CUDA kernels always do the arithmetic; the staged path only adds the copies.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .arrays import DeviceArray


def is_device_leaf(leaf: Any) -> bool:
    return isinstance(leaf, DeviceArray) or (isinstance(leaf, torch.Tensor) and leaf.is_cuda)


def is_array_leaf(leaf: Any) -> bool:
    """What the reference's ``isinstance(leaf, Array)`` asserts accept, widened to device leaves."""
    if is_device_leaf(leaf) or isinstance(leaf, (np.ndarray, torch.Tensor)):
        return True
    return hasattr(leaf, "__array__") and hasattr(leaf, "shape")


def batch_rank(leaf: Any, feature_rank: int = 0) -> int:
    """Number of leading batch dimensions of a leaf."""
    bs = getattr(leaf, "batch_shape", None)
    if bs is not None:
        return len(bs)
    return len(leaf.shape) - feature_rank


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("parllel_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Staging:
    """Resolves leaves to CUDA tensors for one transform call and writes host leaves back."""

    def __init__(self, device: torch.device | None = None) -> None:
        self.device = device
        self.staged = False  # True once any host leaf had to be copied to the device
        self._writebacks: list[tuple[torch.Tensor, torch.Tensor]] = []

    def _adopt_device(self, t: torch.Tensor) -> None:
        if self.device is None:
            self.device = t.device

    def view(self, leaf: Any, writeback: bool = False) -> torch.Tensor:
        if isinstance(leaf, DeviceArray):
            t = leaf.tensor
            self._adopt_device(t)
            return t
        if isinstance(leaf, torch.Tensor) and leaf.is_cuda:
            self._adopt_device(leaf)
            return leaf
        if self.device is None:
            self.device = default_device()
        # zero-copy torch view of the caller's host buffer (pinned memory is recognised by torch
        # through cudaPointerGetAttributes, in which case the copies below are direct DMA)
        self.staged = True
        host = leaf.detach() if isinstance(leaf, torch.Tensor) else torch.from_numpy(np.asarray(leaf))
        dev = torch.empty(host.shape, dtype=host.dtype, device=self.device)
        dev.copy_(host, non_blocking=True)
        if writeback:
            self._writebacks.append((host, dev))
        return dev

    def finish(self) -> None:
        """Write results back into the caller's host buffers (D2H) and wait for them."""
        if not self._writebacks:
            return
        for host, dev in self._writebacks:
            host.copy_(dev, non_blocking=True)
        self._writebacks.clear()
        torch.cuda.current_stream(self.device).synchronize()


def previous_row(leaf: Any, staging: Staging) -> torch.Tensor | None:
    """The row just before the visible window (``leaf[-1]`` in parllel's padding semantics,
    norm_rewards.py:92,95) as a CUDA tensor, or None when the leaf has no padding."""
    if isinstance(leaf, DeviceArray):
        return leaf[-1] if leaf.padding >= 1 else None
    if isinstance(leaf, (torch.Tensor, np.ndarray)):
        return None
    if getattr(leaf, "padding", 0) >= 1:
        row = np.asarray(leaf[-1])
        return torch.from_numpy(np.ascontiguousarray(row)).to(staging.device or default_device())
    return None


def rows_2d(t: torch.Tensor):
    """(data_ptr, row stride in elements, row length) of a tensor seen as [T, row...]:
    the leading axis may be strided (a window into a padded base), every trailing axis
    must be dense -- the views parllel's ``Array`` hands to numba (array.py:420-431)."""
    shape, stride = tuple(t.shape), tuple(t.stride())
    inner = 1
    for size, st in zip(reversed(shape[1:]), reversed(stride[1:])):
        if size != 1 and st != inner:
            raise ValueError("trailing axes of a leaf must be contiguous")
        inner *= size
    ld = stride[0] if (len(shape) >= 1 and shape[0] > 1) else inner
    if ld < inner:
        raise ValueError("overlapping rows are not supported")
    return t.data_ptr(), int(max(ld, 1)), int(inner)


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
