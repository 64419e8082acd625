"""``NormalizeAdvantage`` on the B200 (parllel/transforms/norm_advantage.py:12-56)."""
from __future__ import annotations

import torch

from .. import _lib
from .._leaves import Staging, batch_rank, is_array_leaf, rows_2d, stream_ptr
from .base import Transform

EPSILON = 1e-6


class NormalizeAdvantage(Transform):
    """Batch-normalises ``advantage``: subtract the batch mean, divide by (std + 1e-6).  With a
    ``valid`` field the statistics use valid steps only; all steps are normalised.  With a
    trailing agent axis (``ndim > 2``) statistics are per agent (norm_advantage.py:44-49).

    :param sample_tree: the tree that will be passed to ``__call__``
    """

    def __init__(self, sample_tree) -> None:
        self.only_valid = "valid" in sample_tree
        self.multiagent = len(sample_tree["advantage"].shape) > 2
        self._moments = None
        self._ws = None

    def __call__(self, sample_tree):
        advantage = sample_tree["advantage"]
        assert is_array_leaf(advantage)
        assert batch_rank(advantage, len(advantage.shape) - 2) == 2
        st = Staging()
        adv = st.view(advantage, writeback=True)
        valid = st.view(sample_tree["valid"]) if self.only_valid else None
        self._normalize(adv, valid, st.device)
        st.finish()
        return sample_tree

    def _exchange(self, moments: torch.Tensor, inner: int) -> None:
        """Hook for B-sharded operation (parllel_b200.distributed): single GPU = nothing to exchange."""

    def _normalize(self, adv: torch.Tensor, valid, device) -> None:
        lib = _lib.load()
        s = stream_ptr(device)
        ptr, ld, cols = rows_2d(adv)
        T = adv.shape[0]
        inner = adv.shape[2] if (self.multiagent and adv.dim() > 2) else 1
        if self._moments is None or self._moments.numel() < 3 * inner or self._moments.device != device:
            self._moments = torch.zeros(3 * max(inner, 1), dtype=torch.float64, device=device)
        if inner == 1:
            if self._ws is None or self._ws.device != device:
                self._ws = torch.zeros(lib.plb_batch_moments_workspace_bytes(), dtype=torch.uint8, device=device)
            vptr, vld = (0, 0) if valid is None else rows_2d(valid)[:2]
            _lib.check(lib.plb_batch_moments_f32(ptr, ld, T, cols, vptr, vld, self._moments.data_ptr(),
                                                 self._ws.data_ptr(), self._ws.numel(), s), "plb_batch_moments_f32")
        else:
            # per-agent statistics over all (t, b): column moments of the [T*B, N] view
            flat = adv.reshape(-1, inner)
            rows = flat.shape[0]
            row_valid = valid.reshape(-1).contiguous() if valid is not None else None
            nbytes = lib.plb_column_moments_workspace_bytes(rows, inner)
            if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
                self._ws = torch.zeros(max(nbytes, 64), dtype=torch.uint8, device=device)
            cnt = torch.empty(1, dtype=torch.float64, device=device)
            mean = torch.empty(inner, dtype=torch.float64, device=device)
            m2 = torch.empty(inner, dtype=torch.float64, device=device)
            fp, fld, fcols = rows_2d(flat)
            _lib.check(lib.plb_column_moments_f32(fp, fld, rows, fcols, 0 if row_valid is None else row_valid.data_ptr(),
                                                  cnt.data_ptr(), mean.data_ptr(), m2.data_ptr(),
                                                  self._ws.data_ptr(), self._ws.numel(), s), "plb_column_moments_f32")
            self._moments[: 3 * inner] = torch.stack([cnt.expand(inner), mean, m2], dim=1).reshape(-1)
        self._exchange(self._moments, inner)
        _lib.check(lib.plb_normalize_advantage_f32(ptr, ld, T, cols, inner, self._moments.data_ptr(), EPSILON, s),
                   "plb_normalize_advantage_f32")
