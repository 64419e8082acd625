"""``ClipRewards`` on the B200 (parllel/transforms/clip_rewards.py:9-38)."""
from __future__ import annotations

from .. import _lib
from .._leaves import Staging, is_array_leaf, rows_2d, stream_ptr
from .transform import Transform


class ClipRewards(Transform):
    """Clips rewards to ``[reward_min, reward_max]``; either bound may be ``None``.

    Requires field ``reward`` (any batch shape), modified in place.
    """

    def __init__(self, reward_min: float | None = None, reward_max: float | None = None) -> None:
        self.reward_min = reward_min
        self.reward_max = reward_max
        if not (reward_min is not None or reward_max is not None):
            raise ValueError("Must provide either reward_min or reward_max")

    def __call__(self, sample_tree):
        reward_leaf = sample_tree["reward"]
        assert is_array_leaf(reward_leaf)
        st = Staging()
        reward = st.view(reward_leaf, writeback=True)
        x = reward if reward.dim() >= 2 else reward.reshape(1, -1)
        ptr, ld, cols = rows_2d(x)
        lo = _lib.NAN if self.reward_min is None else float(self.reward_min)
        hi = _lib.NAN if self.reward_max is None else float(self.reward_max)
        rc = _lib.load().plb_scale_clip_f32(ptr, ld, x.shape[0], cols, 0, 0.0, lo, hi, stream_ptr(st.device))
        _lib.check(rc, "ClipRewards")
        st.finish()
        return sample_tree
