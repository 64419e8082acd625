"""``NormalizeObservations`` on the B200 (parllel/transforms/norm_obs.py:15-76)."""
from __future__ import annotations

from collections.abc import Mapping

import numpy as np
import torch

from .. import _lib
from .._leaves import Staging, is_array_leaf, rows_2d, stream_ptr
from .running_mean_std import RunningMeanStd
from .base import Transform

EPSILON = 1e-6


class NormalizeObservations(Transform):
    """Normalises observations with running per-feature statistics: the statistics are updated
    from the step FIRST, then the updated statistics normalise it (norm_obs.py:63-74).  With a
    ``valid`` field only valid environments enter the statistics.

    A step transform: called on ``sample_tree[t]`` with ``observation`` of shape ``[B, *obs_shape]``.
    ``normalize_batch`` applies the same per-step recurrence to a whole ``[T, B, *obs_shape]``
    batch (T dependent steps) without leaving the device.

    :param sample_tree: the tree that will be passed to ``__call__``
    :param obs_shape: shape of one observation
    :param initial_count: seed the running statistics with this many N(0,1) samples (>= 1)
    """

    def __init__(self, sample_tree, obs_shape: tuple, initial_count: float | None = None, device=None,
                 clip: float | None = None) -> None:
        if isinstance(sample_tree["observation"], Mapping):
            raise NotImplementedError("Dictionary observations not supported.")
        self.only_valid = "valid" in sample_tree
        if initial_count is not None and initial_count < 1.0:
            raise ValueError("Initial count must be at least 1")
        self.obs_shape = tuple(obs_shape)
        if initial_count is not None:
            self.obs_statistics = RunningMeanStd(shape=self.obs_shape, initial_count=initial_count, device=device)
        else:
            self.obs_statistics = RunningMeanStd(shape=self.obs_shape, device=device)
        self._ws = None
        # EXTENSION (not in the reference): clamp the normalised observation to [-clip, clip] in the same pass
        self.clip = clip

    def state_dict(self) -> dict:
        return {"obs_statistics": self.obs_statistics.state_dict()}

    def load_state_dict(self, state: dict) -> None:
        self.obs_statistics.load_state_dict(state["obs_statistics"])

    def __call__(self, sample_tree):
        step_obs = sample_tree["observation"]
        assert is_array_leaf(step_obs)
        assert len(step_obs.shape) - len(self.obs_shape) == 1
        st = Staging(self.obs_statistics.device)
        obs = st.view(step_obs, writeback=True)
        valid = st.view(sample_tree["valid"]) if self.only_valid else None
        self._step(obs, valid, st.device)
        st.finish()
        return sample_tree

    def normalize_batch(self, observation, valid=None):
        """Apply the step transform to every ``t`` of a ``[T, B, *obs_shape]`` device batch, in order."""
        st = Staging(self.obs_statistics.device)
        obs = st.view(observation, writeback=True)
        v = st.view(valid) if valid is not None else None
        for t in range(obs.shape[0]):
            self._step(obs[t], None if v is None else v[t], st.device)
        st.finish()
        return observation

    def _step(self, obs: torch.Tensor, valid, device) -> None:
        if self.obs_statistics.sharded:
            # B-sharded: local column moments -> all-gather + merge -> update state -> local normalise
            self.obs_statistics.update(obs, valid=None if valid is None else valid.reshape(-1).contiguous())
            B = obs.shape[0]
            x = obs.reshape(B, -1)
            ptr, ld, cols = rows_2d(x)
            stats = self.obs_statistics
            lo = _lib.NAN if self.clip is None else -float(self.clip)
            hi = _lib.NAN if self.clip is None else float(self.clip)
            _lib.check(_lib.load().plb_normalize_columns_clip_f32(ptr, ld, B, cols, stats.mean_tensor.data_ptr(),
                                                                  stats.var_tensor.data_ptr(), EPSILON, lo, hi,
                                                                  stream_ptr(device)),
                       "plb_normalize_columns_clip_f32")
            return
        lib = _lib.load()
        B = obs.shape[0]
        F = int(np.prod(self.obs_shape, dtype=np.int64)) if self.obs_shape else 1
        x = obs.reshape(B, F) if obs.dim() != 2 else obs
        assert x.data_ptr() == obs.data_ptr(), "observation step must be a dense [B, *obs_shape] view"
        ptr, ld, cols = rows_2d(x)
        assert cols == F
        nbytes = lib.plb_normalize_observations_workspace_bytes(B, F)
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.zeros(max(nbytes, 64), dtype=torch.uint8, device=device)
        stats = self.obs_statistics
        vptr = 0
        if valid is not None:
            valid = valid.reshape(-1).contiguous()
            vptr = valid.data_ptr()
        rc = lib.plb_normalize_observations_step_f32(
            ptr, ld, B, F, vptr, stats.mean_tensor.data_ptr(), stats.var_tensor.data_ptr(),
            stats.count_tensor.data_ptr(), EPSILON, _lib.MOMENTS_F32_FLOW,
            _lib.NAN if self.clip is None else -float(self.clip), _lib.NAN if self.clip is None else float(self.clip),
            self._ws.data_ptr(), self._ws.numel(), stream_ptr(device))
        _lib.check(rc, "NormalizeObservations")
