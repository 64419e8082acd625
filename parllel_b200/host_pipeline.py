"""Host-side helpers for the drop-in (host leaves) path.

``PinnedBatch`` allocates the leaves of a sample tree as numpy arrays backed by page-locked
memory, so that the H2D / D2H copies a transform performs on host leaves are direct DMA
transfers.  The reference allocates its sample tree once (parllel/patterns.py:104-235); this is
the same allocation made DMA-friendly, nothing else changes for the caller.
"""
from __future__ import annotations

import numpy as np
import torch

from .tree import ArrayDict


def pinned_like(array: np.ndarray) -> np.ndarray:
    """A page-locked numpy array with the contents of ``array``."""
    t = torch.empty(array.shape, dtype=torch.from_numpy(np.empty(0, array.dtype)).dtype).pin_memory()
    out = t.numpy()
    out[...] = array
    out_holder = _PinnedHolder(out, t)
    return out_holder.array


class _PinnedHolder:
    """Keeps the owning pinned tensor alive for as long as the numpy view lives."""
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
