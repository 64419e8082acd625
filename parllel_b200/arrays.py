"""``DeviceArray``: the HBM-resident counterpart of parllel's ``Array`` storage layout.

Honours the layout contract of parllel/arrays/array.py (not its indexing algebra):

* base buffer ``[full_size + 2*padding, *batch_shape[1:], *feature_shape]``, C-contiguous
  (array.py:133), here one ``torch`` CUDA tensor so every leaf is a single HBM slab;
* the visible window is ``T = batch_shape[0]`` rows starting at ``padding + offset``;
* leading indices outside ``[0, T)`` address the padding rows, they do NOT wrap
  (``neg_from_end=False``, array.py:302,322): ``arr[-1]`` is the row before the window;
* ``rotate()`` slides the window by ``T`` and, on wrap-around, copies the last ``padding``
  rows + trailing padding to the front (array.py:394-418);
* ``.full`` exposes the whole ring without padding (array.py:330-346).

Only leading-axis ints/slices (plus ordinary trailing indices) are supported -- that
is all the transforms, samplers' ``tree[t]`` and the replay ring need.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

_NP_TO_TORCH = {
    np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
    np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
    np.dtype(np.uint8): torch.uint8, np.dtype(np.bool_): torch.bool,
    np.dtype(np.int16): torch.int16, np.dtype(np.int8): torch.int8, np.dtype(np.float16): torch.float16,
}


def _torch_dtype(dtype) -> torch.dtype:
    if isinstance(dtype, torch.dtype):
        return dtype
    return _NP_TO_TORCH[np.dtype(dtype)]


class DeviceArray:
    def __init__(self, batch_shape: tuple, dtype, *, feature_shape: tuple = (), padding: int = 0,
                 full_size: int | None = None, device: Any = "cuda") -> None:
        if not batch_shape:
            raise ValueError("Non-empty batch_shape required.")
        if padding < 0:
            raise ValueError("Padding must be non-negative.")
        if padding > batch_shape[0]:
            raise ValueError(f"Padding ({padding}) cannot be greater than leading dimension {batch_shape[0]}.")
        T = int(batch_shape[0])
        if full_size is None:
            full_size = T
        if full_size < T:
            raise ValueError(f"Full size ({full_size}) cannot be less than leading dimension {T}.")
        if full_size % T != 0:
            raise ValueError(f"Full size ({full_size}) must be evenly divided by leading dimension {T}.")
        self.dtype = _torch_dtype(dtype)
        self.padding = int(padding)
        self.full_size = int(full_size)
        self._T = T
        self._n_batch_dims = len(batch_shape)
        self._base = torch.zeros((self.full_size + 2 * self.padding,) + tuple(batch_shape[1:]) + tuple(feature_shape),
                                 dtype=self.dtype, device=device)
        self._start = self.padding  # first base row of the visible window
        self._rotatable = True

    # -- construction helpers ---------------------------------------------------------
    @classmethod
    def from_host(cls, array, device: Any = "cuda") -> "DeviceArray":
        """Mirror a parllel ``Array`` (same padding / full_size / window position) on the device."""
        base = np.asarray(array._base_array)
        n_batch = array._base_batch_dims
        full = int(array.full_size)
        pad = int(array.padding)
        T = int(array.shape[0])
        out = cls((T,) + tuple(base.shape[1:n_batch]), base.dtype, feature_shape=tuple(base.shape[n_batch:]),
                  padding=pad, full_size=full, device=device)
        out._base.copy_(torch.from_numpy(np.ascontiguousarray(base)))
        lead = array._current_location[0]
        out._start = int(lead.start)
        return out

    # -- geometry ---------------------------------------------------------------------
    @property
    def shape(self) -> tuple:
        return (self._T,) + tuple(self._base.shape[1:])

    @property
    def batch_shape(self) -> tuple:
        return self.shape[: self._n_batch_dims]

    @property
    def n_batch_dims(self) -> int:
        return self._n_batch_dims

    @property
    def first(self) -> int:
        return 0

    @property
    def last(self) -> int:
        return self._T - 1

    @property
    def device(self):
        return self._base.device

    @property
    def base(self) -> torch.Tensor:
        return self._base

    @property
    def tensor(self) -> torch.Tensor:
        """The visible ``[T, ...]`` window as a zero-copy view."""
        return self._base[self._start: self._start + self._T]

    @property
    def full(self) -> "DeviceArray":
        view = object.__new__(DeviceArray)
        view.__dict__.update(self.__dict__)
        view._T = self.full_size
        view._start = self.padding
        view._rotatable = False
        return view

    # -- indexing (leading axis relative to the window; out-of-window = padding) --------
    def _lead(self, idx):
        lo, hi = -self._start, self._base.shape[0] - self._start  # addressable window-relative range
        if isinstance(idx, (int, np.integer)):
            i = int(idx)
            if not lo <= i < hi:
                raise IndexError(f"index {i} outside [{lo}, {hi}) (window of {self._T} rows + padding)")
            return self._start + i
        if isinstance(idx, slice):
            step = 1 if idx.step is None else idx.step
            if step <= 0:
                raise IndexError("only positive slice steps are supported")
            a = 0 if idx.start is None else int(idx.start)
            b = self._T if idx.stop is None else int(idx.stop)
            if a < lo or b > hi:
                raise IndexError(f"slice {idx} outside [{lo}, {hi})")
            return slice(self._start + a, self._start + max(a, b), step)
        if idx is Ellipsis:
            return slice(self._start, self._start + self._T)
        raise IndexError(f"unsupported leading index {idx!r}")

    def _location(self, key):
        if isinstance(key, tuple):
            if len(key) == 0:
                return (self._lead(Ellipsis),)
            if key[0] is Ellipsis:
                return (self._lead(Ellipsis),) + tuple(key)
            return (self._lead(key[0]),) + tuple(key[1:])
        return (self._lead(key),)

    def __getitem__(self, key) -> torch.Tensor:
        return self._base[self._location(key)]

    def __setitem__(self, key, value) -> None:
        if isinstance(value, np.ndarray):
            value = torch.from_numpy(np.ascontiguousarray(value))
        if isinstance(value, torch.Tensor):
            value = value.to(device=self._base.device, dtype=self.dtype)
        self._base[self._location(key)] = value

    # -- ring behaviour ---------------------------------------------------------------
    def rotate(self) -> None:
        if not self._rotatable:
            raise ValueError("rotate() is only allowed on unindexed array.")
        start = self._start + self._T
        if start >= self.full_size:  # same (base-coordinate) test as array.py:402
            start -= self.full_size
            if self.padding:
                p, n = self.padding, self.full_size
                self._base[0: 2 * p].copy_(self._base[n: n + 2 * p].clone())
        self._start = start

    def reset(self) -> None:
        if not self._rotatable:
            raise ValueError("reset() is only allowed on unindexed array.")
        self._start = self.full_size + self.padding - self._T

    # -- host interop -----------------------------------------------------------------
    def to_ndarray(self) -> np.ndarray:
        return self.tensor.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None) -> np.ndarray:
        out = self.to_ndarray()
        return out.astype(dtype, copy=False) if dtype is not None else out

    def __repr__(self) -> str:
        return (f"DeviceArray(shape={self.shape}, dtype={self.dtype}, padding={self.padding}, "
                f"full_size={self.full_size}, device={self._base.device})")

    def close(self) -> None:
        pass
