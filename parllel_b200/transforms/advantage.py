"""``EstimateAdvantage`` on the B200 (parllel/transforms/advantage.py:60-106).

Same constructor, same required tree fields, same in-place contract; the masked
reverse scans run in ``libparllel_b200`` (csrc/plb_scan.cu).
"""
from __future__ import annotations

from .. import _lib
from .._leaves import Staging, batch_rank, is_array_leaf, rows_2d, stream_ptr
from .base import Transform


class EstimateAdvantage(Transform):
    """Adds ``advantage`` and ``return_`` for every sample: GAE(lambda) when ``gae_lambda < 1``,
    empirical discounted return minus value when ``gae_lambda == 1`` (advantage.py:82-89).

    Requires fields ``reward``, ``done``, ``agent_info.value``, ``bootstrap_value``; writes
    ``advantage`` and ``return_``.

    :param discount: discount (gamma)
    :param gae_lambda: lambda of GAE
    :param algo: ``"auto"`` | ``"serial"`` (the reference's per-step float32 rounding,
        bit-identical to the oracle) | ``"chunked"`` (TMA-staged affine scan)
    """

    _ALGOS = {"auto": _lib.ALGO_AUTO, "serial": _lib.ALGO_SERIAL, "chunked": _lib.ALGO_CHUNKED}

    def __init__(self, discount: float, gae_lambda: float, algo: str = "auto") -> None:
        self.discount = discount
        self.gae_lambda = gae_lambda
        self.algo = self._ALGOS[algo]

    def __call__(self, sample_tree):
        advantage = sample_tree["advantage"]
        assert is_array_leaf(advantage)
        value_leaf = sample_tree["agent_info"]["value"]
        reward_leaf = sample_tree["reward"]
        assert batch_rank(reward_leaf) == 2 and batch_rank(advantage, len(advantage.shape) - 2) == 2

        st = Staging()
        reward = st.view(reward_leaf)
        value = st.view(value_leaf)
        done = st.view(sample_tree["done"])
        boot = st.view(sample_tree["bootstrap_value"]).contiguous()
        adv = st.view(advantage, writeback=True)
        ret = st.view(sample_tree["return_"], writeback=True)

        T, B = reward.shape[0], reward.shape[1]
        rp, ldr, nr = rows_2d(reward)
        vp, ldv, nv = rows_2d(value)
        dp, ldd, nd = rows_2d(done)
        ap, lda, na = rows_2d(adv)
        tp, ldt, nt = rows_2d(ret)
        assert nr == B and nd == B and na == nv and nt == nv and nv % max(B, 1) == 0
        inner = nv // max(B, 1)
        assert boot.numel() == nv
        lib = _lib.load()
        s = stream_ptr(st.device)
        if self.gae_lambda == 1.0:
            rc = lib.plb_compute_discount_return_f32(rp, ldr, vp, ldv, dp, ldd, boot.data_ptr(), ap, lda, tp, ldt,
                                                     T, B, inner, float(self.discount), self.algo, s)
        else:
            rc = lib.plb_compute_gae_advantage_f32(rp, ldr, vp, ldv, dp, ldd, boot.data_ptr(), ap, lda, tp, ldt,
                                                   T, B, inner, float(self.discount), float(self.gae_lambda),
                                                   self.algo, s)
        _lib.check(rc, "EstimateAdvantage")
        st.finish()
        return sample_tree
