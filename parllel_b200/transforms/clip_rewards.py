"""Reward clipping on the B200 (counterpart of parllel/transforms/clip_rewards.py:9-38)."""
from __future__ import annotations

from .. import _lib
from .._leaves import Staging, is_array_leaf, rows_2d, stream_ptr
from .base import Transform


class ClipRewards(Transform):
    """Clamp ``reward`` into ``[reward_min, reward_max]`` in place; a bound left at ``None`` is open.
    At least one bound is required (``ValueError`` otherwise, as in the reference).  Any batch shape."""

    reads = ("reward",)
    writes = ("reward",)

    def __init__(self, reward_min: float | None = None, reward_max: float | None = None) -> None:
        if reward_min is None and reward_max is None:
            raise ValueError("Must provide either reward_min or reward_max")
        self.reward_min, self.reward_max = reward_min, reward_max

    def _bounds(self) -> tuple[float, float]:
        nan = _lib.NAN  # NaN = "no bound" in the C ABI
        lo = nan if self.reward_min is None else float(self.reward_min)
        hi = nan if self.reward_max is None else float(self.reward_max)
        return lo, hi

    def __call__(self, sample_tree):
        leaf = sample_tree["reward"]
        assert is_array_leaf(leaf)
        staging = Staging()
        reward = staging.view(leaf, writeback=True)
        flat = reward.reshape(1, -1) if reward.dim() < 2 else reward
        ptr, ld, cols = rows_2d(flat)
        lo, hi = self._bounds()
        rc = _lib.load().plb_scale_clip_f32(ptr, ld, flat.shape[0], cols, 0, 0.0, lo, hi, stream_ptr(staging.device))
        _lib.check(rc, "ClipRewards")
        staging.finish()
        return sample_tree
