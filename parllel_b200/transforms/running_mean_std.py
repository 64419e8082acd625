"""Device-resident ``RunningMeanStd`` (parllel/transforms/running_mean_std.py:35-57).

State ``mean``/``var`` (float64, ``shape``) and ``count`` (float64 scalar) live in HBM so a
transform never synchronises with the host; the ``mean`` / ``var`` / ``count`` properties
copy them out as numpy float64 arrays, which is what callers of the reference read
(``transform.obs_statistics.mean`` ...).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._leaves import default_device, rows_2d, stream_ptr


class RunningMeanStd:
    def __init__(self, shape: tuple, initial_count: float = 1e-4, device=None) -> None:
        self.shape = tuple(shape)
        self.device = torch.device(device) if device is not None else default_device()
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self._n = n
        self.mean_tensor = torch.zeros(n, dtype=torch.float64, device=self.device)
        self.var_tensor = torch.ones(n, dtype=torch.float64, device=self.device)
        self.count_tensor = torch.full((1,), float(initial_count), dtype=torch.float64, device=self.device)
        # scratch for batch moments: scalar -> (count, mean, M2); per column -> count[1], mean[n], m2[n]
        self._moments = torch.zeros(3 if n == 1 else 1 + 2 * n, dtype=torch.float64, device=self.device)
        self._ws = None
        # B-sharded operation (parllel_b200.distributed.shard_statistics): batch moments are all-gathered
        # over this process group and merged in rank order before they update the running state
        self.sharded = False
        self.group = None

    # -- numpy views of the state (D2H copies) ------------------------------------------
    @property
    def mean(self) -> np.ndarray:
        return self.mean_tensor.cpu().numpy().reshape(self.shape)

    @property
    def var(self) -> np.ndarray:
        return self.var_tensor.cpu().numpy().reshape(self.shape)

    @property
    def count(self) -> np.ndarray:
        return self.count_tensor.cpu().numpy().reshape(())

    def load_state(self, mean, var, count) -> None:
        self.mean_tensor.copy_(torch.as_tensor(np.asarray(mean, np.float64).reshape(-1)))
        self.var_tensor.copy_(torch.as_tensor(np.asarray(var, np.float64).reshape(-1)))
        self.count_tensor.fill_(float(count))

    def state_dict(self) -> dict:
        """Checkpointable copy of the running state (numpy float64), for ``load_state_dict``."""
        return {"mean": self.mean.copy(), "var": self.var.copy(), "count": float(self.count)}

    def load_state_dict(self, state: dict) -> None:
        self.load_state(state["mean"], state["var"], state["count"])

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.zeros(max(nbytes, 64), dtype=torch.uint8, device=self.device)
        return self._ws

    # -- updates --------------------------------------------------------------------------
    def update(self, batch: torch.Tensor, valid: torch.Tensor | None = None, f32_flow: bool = True) -> None:
        """``update(batch)`` of the reference: batch moments (np.mean / np.var, :52-53) then the
        float64 Chan merge (:7-32), all on the device.  ``batch`` is ``[rows, *shape]`` (per-feature
        statistics, ``valid`` = per-row mask) or any shape when ``shape == ()`` (scalar statistics,
        ``valid`` = per-element mask)."""
        lib = _lib.load()
        s = stream_ptr(self.device)
        flags = _lib.MOMENTS_F32_FLOW if f32_flow else 0
        if self._n == 1 and self.shape == ():
            x = batch if batch.dim() >= 2 else batch.reshape(1, -1)
            ptr, ld, cols = rows_2d(x)
            rows = x.shape[0]
            vptr, vld = 0, 0
            if valid is not None:
                v = valid if valid.dim() >= 2 else valid.reshape(1, -1)
                vptr, vld, _ = rows_2d(v)
            ws = self._workspace(lib.plb_batch_moments_workspace_bytes())
            _lib.check(lib.plb_batch_moments_f32(ptr, ld, rows, cols, vptr, vld, self._moments.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), s), "plb_batch_moments_f32")
            m = self._moments
            if self.sharded:
                self._exchange(layout=0)
            _lib.check(lib.plb_update_from_moments_f64(m.data_ptr(), m.data_ptr() + 8, m.data_ptr() + 16,
                                                       self.mean_tensor.data_ptr(), self.var_tensor.data_ptr(),
                                                       self.count_tensor.data_ptr(), 1, flags, s),
                       "plb_update_from_moments_f64")
        else:
            x = batch.reshape(batch.shape[0], -1) if batch.dim() != 2 else batch
            ptr, ld, cols = rows_2d(x)
            assert cols == self._n, f"batch feature size {cols} != statistics size {self._n}"
            rows = x.shape[0]
            vptr = valid.data_ptr() if valid is not None else 0
            nbytes = lib.plb_column_moments_workspace_bytes(rows, cols)
            ws = self._workspace(nbytes)
            m = self._moments
            cnt, mean, m2 = m.data_ptr(), m.data_ptr() + 8, m.data_ptr() + 8 * (1 + self._n)
            _lib.check(lib.plb_column_moments_f32(ptr, ld, rows, cols, vptr, cnt, mean, m2,
                                                  ws.data_ptr(), ws.numel(), s), "plb_column_moments_f32")
            if self.sharded:
                self._exchange(layout=1)
            _lib.check(lib.plb_update_from_moments_f64(cnt, mean, m2, self.mean_tensor.data_ptr(),
                                                       self.var_tensor.data_ptr(), self.count_tensor.data_ptr(),
                                                       self._n, flags, s), "plb_update_from_moments_f64")

    def _exchange(self, layout: int) -> None:
        from ..distributed import exchange_moments
        exchange_moments(self._moments, 1 if layout == 0 else self._n, layout, self.group)
