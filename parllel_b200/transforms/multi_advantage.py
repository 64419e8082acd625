"""``EstimateMultiAgentAdvantage`` on the B200 (parllel/transforms/multi_advantage.py:24-147).

Per-agent advantage from a shared (or per-agent) reward and agent-specific value estimates: the same
masked reverse scans as ``EstimateAdvantage`` with a trailing agent axis, broadcast left to right
(parllel/np_utils.py:4-24).  Three problem types, as in the reference:

* single critic   -- value ``[T, B]``, advantage ``[T, B, 1]`` (trailing singleton for broadcasting);
* independent critics -- reward ``[T, B]``, value / advantage ``[T, B, N]``;
* markov game     -- reward is a dict of ``[T, B]`` leaves, stacked in the order of the ``action`` dict.
  (The reference builds that stack with ``np.stack(<generator>)``, multi_advantage.py:118-121, which
  numpy >= 1.24 rejects; the evident intent is implemented and pinned against the oracle only.)
"""
from __future__ import annotations

from collections.abc import Mapping
from enum import Enum

import torch

from .. import _lib
from .._leaves import Staging, is_array_leaf, rows_2d, stream_ptr
from .base import Transform


class ProblemType(Enum):
    single_critic = 0
    independent_critics = 1
    markov_game = 3


class EstimateMultiAgentAdvantage(Transform):
    """Requires ``reward``, ``done``, ``action``, ``agent_info.value``, ``bootstrap_value``; writes
    ``advantage`` and ``return_``.

    :param sample_tree: the tree that will be passed to ``__call__``
    :param discount: discount (gamma)
    :param gae_lambda: lambda of GAE (1.0 = empirical discounted return)
    """

    def __init__(self, sample_tree, discount: float, gae_lambda: float) -> None:
        self._discount = discount
        self._lambda = gae_lambda
        reward = sample_tree["reward"]
        value = sample_tree["agent_info"]["value"]
        advantage = sample_tree["advantage"]
        if len(value.shape) > 2:
            self.problem_type = ProblemType.markov_game if isinstance(reward, Mapping) else ProblemType.independent_critics
            if tuple(advantage.shape) != tuple(value.shape):
                raise ValueError("Advantage and Value must have the same shape")
        else:
            self.problem_type = ProblemType.single_critic
            if tuple(advantage.shape) != tuple(value.shape) + (1,):
                raise ValueError(
                    "Advantage must match shape of Value except for added trailing singleton dimension.")

    def __call__(self, sample_tree):
        done_leaf = sample_tree["done"]
        assert is_array_leaf(done_leaf)
        assert len(done_leaf.shape) == 2
        st = Staging()
        done = st.view(done_leaf)
        value = st.view(sample_tree["agent_info"]["value"])
        boot = st.view(sample_tree["bootstrap_value"]).contiguous()
        adv = st.view(sample_tree["advantage"], writeback=True)
        ret = st.view(sample_tree["return_"], writeback=True)
        T, B = done.shape
        if self.problem_type is ProblemType.markov_game:
            reward_tree, action = sample_tree["reward"], sample_tree["action"]
            assert isinstance(reward_tree, Mapping) and isinstance(action, Mapping)
            # one column per (env, agent): reward stacked in the order of the action dict, done repeated
            reward = torch.stack([st.view(reward_tree[name]) for name in action], dim=-1).contiguous()
            N = reward.shape[-1]
            done = done.unsqueeze(-1).expand(T, B, N).contiguous()
            reward, done = reward.reshape(T, B * N), done.reshape(T, B * N)
            cols, inner = B * N, 1
        else:
            reward_leaf = sample_tree["reward"]
            assert is_array_leaf(reward_leaf)
            reward = st.view(reward_leaf)
            cols = B
            inner = value.numel() // max(T * B, 1)
        rp, ldr, _ = rows_2d(reward)
        dp, ldd, _ = rows_2d(done)
        vp, ldv, nv = rows_2d(value)
        ap, lda, na = rows_2d(adv)
        tp, ldt, nt = rows_2d(ret)
        assert na == nv == cols * inner and nt == nv and boot.numel() == nv
        lib = _lib.load()
        s = stream_ptr(st.device)
        if self._lambda == 1.0:
            rc = lib.plb_compute_discount_return_f32(rp, ldr, vp, ldv, dp, ldd, boot.data_ptr(), ap, lda, tp, ldt,
                                                     T, cols, inner, float(self._discount), _lib.ALGO_AUTO, s)
        else:
            rc = lib.plb_compute_gae_advantage_f32(rp, ldr, vp, ldv, dp, ldd, boot.data_ptr(), ap, lda, tp, ldt,
                                                   T, cols, inner, float(self._discount), float(self._lambda),
                                                   _lib.ALGO_AUTO, s)
        _lib.check(rc, "EstimateMultiAgentAdvantage")
        st.finish()
        return sample_tree
