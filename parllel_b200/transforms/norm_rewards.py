"""``NormalizeRewards`` on the B200 (parllel/transforms/norm_rewards.py:36-116)."""
from __future__ import annotations

from collections.abc import Mapping

import torch

from .. import _lib
from .._leaves import Staging, batch_rank, is_array_leaf, previous_row, rows_2d, stream_ptr
from .running_mean_std import RunningMeanStd
from .base import Transform

EPSILON = 1e-6


class NormalizeRewards(Transform):
    """Divides rewards by the running standard deviation of the past discounted return and
    writes ``past_return`` as a side effect.  With a ``valid`` field only valid steps enter
    the statistics.  No mean subtraction (norm_rewards.py:114).

    Requires fields ``reward``, ``done``, ``[valid]``; writes ``past_return`` (+ ``reward``).

    The carry rows ``past_return[-1]`` / ``done[-1]`` (norm_rewards.py:92,95) are read from the
    padding of the leaves when they have one (parllel ``Array`` / ``DeviceArray`` with
    ``padding >= 1``, rotated by the sampler).  Plain tensors have no padding; the transform
    then keeps the last row of the previous call itself, which is what ``rotate()`` would
    have copied there.

    :param sample_tree: the tree that will be passed to ``__call__``
    :param discount: discount (gamma)
    :param initial_count: seed the running statistics with this many N(0,1) samples (>= 1)
    """

    _ALGOS = {"auto": _lib.ALGO_AUTO, "serial": _lib.ALGO_SERIAL, "chunked": _lib.ALGO_CHUNKED}

    def __init__(self, sample_tree, discount: float, initial_count: float | None = None,
                 algo: str = "auto", device=None) -> None:
        if isinstance(sample_tree["reward"], Mapping):
            raise NotImplementedError("Not implemented for markov games, where rewards are agent-specific.")
        self.only_valid = "valid" in sample_tree
        self.discount = discount
        if initial_count is not None and initial_count < 1.0:
            raise ValueError("Initial count must be at least 1")
        if initial_count is not None:
            self.return_statistics = RunningMeanStd(shape=(), initial_count=initial_count, device=device)
        else:
            self.return_statistics = RunningMeanStd(shape=(), device=device)
        self.algo = self._ALGOS[algo]
        self._carry_return = None  # used only for leaves without padding
        self._carry_done = None
        self._fused = None

    # -- checkpoint / resume (the reference keeps this state in plain attributes that pickle with it) ----
    def state_dict(self) -> dict:
        state = {"return_statistics": self.return_statistics.state_dict()}
        if self._carry_return is not None:
            state["carry_return"] = self._carry_return.cpu().numpy().copy()
        if self._carry_done is not None:
            state["carry_done"] = self._carry_done.cpu().numpy().copy()
        return state

    def load_state_dict(self, state: dict) -> None:
        self.return_statistics.load_state_dict(state["return_statistics"])
        dev = self.return_statistics.device
        if "carry_return" in state:
            self._carry_return = torch.as_tensor(state["carry_return"], dtype=torch.float32).to(dev)
        if "carry_done" in state:
            self._carry_done = torch.as_tensor(state["carry_done"], dtype=torch.bool).to(dev)

    def __call__(self, sample_tree):
        reward_leaf = sample_tree["reward"]
        assert is_array_leaf(reward_leaf)
        assert batch_rank(reward_leaf) in (1, 2)
        past_leaf, done_leaf = sample_tree["past_return"], sample_tree["done"]
        if (self.algo != _lib.ALGO_SERIAL and batch_rank(reward_leaf) == 2 and not self.return_statistics.sharded
                and len(reward_leaf.shape) == 2):
            # [T, B] batches take the fused chain's route: the past-return scan accumulates the moments and
            # folds them into the running state in its tail, then one scaling pass (2 launches instead of 4)
            if self._fused is None:
                from .pipeline import FusedBatchTransforms
                self._fused = FusedBatchTransforms([self])
            return self._fused(sample_tree)

        st = Staging(self.return_statistics.device)
        reward = st.view(reward_leaf, writeback=True)
        past = st.view(past_leaf, writeback=True)
        done = st.view(done_leaf)
        valid = st.view(sample_tree["valid"]) if self.only_valid else None
        prev_ret = previous_row(past_leaf, st)
        prev_done = previous_row(done_leaf, st)

        one_d = reward.dim() == 1  # batch_shape of length 1: a single trajectory [T]
        if one_d:
            reward2, past2, done2 = reward.unsqueeze(1), past.unsqueeze(1), done.unsqueeze(1)
            valid2 = valid.unsqueeze(1) if valid is not None else None
        else:
            reward2, past2, done2, valid2 = reward, past, done, valid
        T, B = reward2.shape
        dev = st.device
        if prev_ret is None:
            if self._carry_return is None:
                self._carry_return = torch.zeros(B, dtype=torch.float32, device=dev)
            prev_ret = self._carry_return
        if prev_done is None:
            if self._carry_done is None:
                self._carry_done = torch.zeros(B, dtype=torch.bool, device=dev)
            prev_done = self._carry_done
        prev_ret = prev_ret.reshape(-1).contiguous()
        prev_done = prev_done.reshape(-1).contiguous()

        lib = _lib.load()
        s = stream_ptr(dev)
        rp, ldr, _ = rows_2d(reward2)
        dp, ldd, _ = rows_2d(done2)
        pp, ldp, _ = rows_2d(past2)
        _lib.check(lib.plb_compute_past_discount_return_f32(rp, ldr, dp, ldd, prev_ret.data_ptr(), prev_done.data_ptr(),
                                                            pp, ldp, T, B, float(self.discount), self.algo, s),
                   "plb_compute_past_discount_return_f32")
        # update statistics of the discounted return, then rescale the rewards
        self.return_statistics.update(past2, valid=valid2)
        _lib.check(lib.plb_scale_clip_f32(rp, ldr, T, B, self.return_statistics.var_tensor.data_ptr(), EPSILON,
                                          _lib.NAN, _lib.NAN, s), "plb_scale_clip_f32")
        if T > 0:
            if self._carry_return is not None:
                self._carry_return.copy_(past2[T - 1])
            if self._carry_done is not None:
                self._carry_done.copy_(done2[T - 1])
        st.finish()
        return sample_tree
