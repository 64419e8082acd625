"""GPU parity of the resident advantage kernel: EstimateAdvantage + NormalizeAdvantage as one cooperative
launch (advantage parked in tensor memory, statistics traded between all CTAs through the grid-exchange
records), against the CPU oracle (advantage.py:36-57 then norm_advantage.py:30-56).

Tolerance: |got - ref| <= 1e-5 * max(|ref|, rms(ref)) (tests/util.py; BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.util import assert_within_tolerance, make_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _resident_on_one_gpu():
    """The library's default policy takes the resident kernel only when B is sharded over several GPUs (on one
    GPU the two-kernel route is faster); these tests force it on the single GPU they run on."""
    from parllel_b200 import _lib
    lib = _lib.load()
    lib.plb_set_tuning(12, 2)
    yield
    lib.plb_set_tuning(12, 1)


def _tree(batch, dev="cuda", extra=()):
    from parllel_b200 import ArrayDict
    t = {k: torch.from_numpy(v).to(dev) for k, v in batch.items()}
    d = {"reward": t["reward"], "done": t["done"], "agent_info": {"value": t["value"]},
         "bootstrap_value": t["bootstrap_value"],
         "advantage": torch.full_like(t["value"], float("nan")),
         "return_": torch.full_like(t["value"], float("nan"))}
    for name in extra:
        d[name] = torch.full_like(t["value"], float("nan"))
    if "valid" in t:
        d["valid"] = t["valid"]
    return ArrayDict(d)


def _oracle_step(batch, gamma, lam, valid=None):
    adv, ret = np.empty_like(batch["value"]), np.empty_like(batch["value"])
    fn = O.compute_discount_return if lam == 1.0 else O.compute_gae_advantage
    fn(batch["reward"], batch["value"], batch["done"], batch["bootstrap_value"], gamma, lam, adv, ret)
    O.normalize_advantage(adv, valid)
    return adv, ret


def _fused(tree, lam, graph=False, launch=None):
    from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage
    return FusedBatchTransforms([EstimateAdvantage(0.99, lam), NormalizeAdvantage(tree)], use_cuda_graph=graph, launch=launch)


# (T, B): one tile; ragged T below / above one tile; T <= 64 (64-row tiles); many strips; a ragged last
# strip (B % 32 != 0); two strips per CTA (the C4/8 shard); the largest strip tensor memory holds (16 tiles)
SHAPES = [(1, 16), (5, 48), (63, 32), (64, 64), (65, 80), (128, 4096), (300, 1040), (256, 8192), (2048, 32),
          (1024, 4096)]


@pytest.mark.parametrize("lam", [0.95, 1.0])
@pytest.mark.parametrize("T,B", SHAPES)
def test_resident_step_vs_oracle(T, B, lam):
    from parllel_b200 import _lib
    batch = make_batch(np.random.default_rng(T * 131 + B), T, B, 0.02)
    adv_ref, ret_ref = _oracle_step(batch, 0.99, lam)
    tree = _tree(batch)
    before = _lib.launch_counters()
    assert _fused(tree, lam)(tree) is tree
    after = _lib.launch_counters()
    assert after[1] - before[1] == 1 and after[0] - before[0] == 1, "expected exactly one (resident) launch"
    assert_within_tolerance(tree["return_"].cpu().numpy(), ret_ref, what=f"return_ T={T} B={B}")
    assert_within_tolerance(tree["advantage"].cpu().numpy(), adv_ref, what=f"advantage T={T} B={B}")


@pytest.mark.parametrize("T,B", [(65, 80), (300, 1040), (1024, 4096)])
def test_resident_equals_two_kernel_route(T, B):
    """Same arithmetic per element; only the order in which the partial moments are merged differs, so the
    two routes agree to float32 rounding of (mean, std) -- far inside the parity tolerance."""
    from parllel_b200 import _lib
    lib = _lib.load()
    batch = make_batch(np.random.default_rng(7), T, B, 0.01)
    a, b = _tree(batch), _tree(batch)
    _fused(a, 0.95)(a)
    try:
        lib.plb_set_tuning(12, 0)
        before = _lib.launch_counters()
        _fused(b, 0.95)(b)
        assert _lib.launch_counters()[1] == before[1], "resident kernel ran although switched off"
    finally:
        lib.plb_set_tuning(12, 2)
    assert torch.equal(a["return_"], b["return_"])
    np.testing.assert_allclose(a["advantage"].cpu().numpy(), b["advantage"].cpu().numpy(), rtol=3e-7, atol=3e-7)


@pytest.mark.parametrize("T,B", [(65, 80), (1024, 4096)])
def test_resident_plain_store_variant_is_bit_identical(T, B):
    """Tuning key 5 bit 2 sends the final advantage through plain stores instead of per-warp TMA stores."""
    from parllel_b200 import _lib
    lib = _lib.load()
    batch = make_batch(np.random.default_rng(8), T, B, 0.01)
    a, b = _tree(batch), _tree(batch)
    _fused(a, 0.95)(a)
    try:
        lib.plb_set_tuning(5, 4)
        _fused(b, 0.95)(b)
    finally:
        lib.plb_set_tuning(5, 0)
    assert torch.equal(a["advantage"], b["advantage"]) and torch.equal(a["return_"], b["return_"])


def test_resident_with_valid_mask_vs_oracle():
    T, B = 200, 512
    rng = np.random.default_rng(11)
    batch = make_batch(rng, T, B, 0.02)
    valid = np.ones((T, B), bool)
    for t in range(T - 1):
        valid[t + 1] = valid[t] & ~batch["done"][t]
    adv_ref, ret_ref = _oracle_step(batch, 0.99, 0.95, valid)
    tree = _tree(dict(batch, valid=valid))
    _fused(tree, 0.95)(tree)
    assert_within_tolerance(tree["advantage"].cpu().numpy(), adv_ref, what="masked advantage")
    assert_within_tolerance(tree["return_"].cpu().numpy(), ret_ref, what="masked return_")


@pytest.mark.parametrize("launch", ["graph", "direct"])
def test_resident_replays_through_a_graph_and_epochs_advance(launch):
    """Many launches on one buffer set (eager, captured, replayed -- or the bound descriptor launched directly):
    every launch must see only its own epoch's records.  Inputs change between launches, so a stale record would show."""
    from parllel_b200 import _lib
    T, B = 256, 2048
    rng = np.random.default_rng(5)
    batch = make_batch(rng, T, B, 0.01)
    tree = _tree(batch)
    fused = _fused(tree, 0.95, graph=True, launch=launch)
    for it in range(7):
        scale = np.float32(1.0 + it)
        cur = dict(batch, reward=batch["reward"] * scale)
        tree["reward"].copy_(torch.from_numpy(cur["reward"]))
        fused(tree)
        adv_ref, ret_ref = _oracle_step(cur, 0.99, 0.95)
        assert_within_tolerance(tree["advantage"].cpu().numpy(), adv_ref, what=f"launch {it} advantage")
        assert_within_tolerance(tree["return_"].cpu().numpy(), ret_ref, what=f"launch {it} return_")
    assert fused._gx is not None and fused._gx.status() == 0


@pytest.mark.parametrize("launch", ["graph", "direct"])
def test_swapping_a_leaf_in_the_same_tree_does_not_replay_a_stale_graph(launch):
    """the replay fast path must look at the leaves the tree holds NOW."""
    T, B = 128, 256
    batch = make_batch(np.random.default_rng(9), T, B, 0.01)
    tree = _tree(batch)
    fused = _fused(tree, 0.95, graph=True, launch=launch)
    for _ in range(4):  # eager, capture, bound replays
        fused(tree)
    other = dict(batch, reward=batch["reward"] * np.float32(3.0))
    tree["reward"] = torch.from_numpy(other["reward"]).cuda()  # a NEW tensor under the same key
    tree["advantage"] = torch.full((T, B), float("nan"), device="cuda")
    fused(tree)
    adv_ref, _ = _oracle_step(other, 0.99, 0.95)
    assert_within_tolerance(tree["advantage"].cpu().numpy(), adv_ref, what="advantage after leaf swap")


def test_full_chain_with_resident_tail_vs_oracle():
    """NormalizeRewards -> ClipRewards -> EstimateAdvantage -> NormalizeAdvantage at a multi-strip shape: the
    advantage scan scales/clips the reward on its load path AND keeps the advantage resident."""
    from parllel_b200 import _lib
    from parllel_b200.transforms import (ClipRewards, EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage,
                                         NormalizeRewards)
    T, B = 384, 4096
    batch = make_batch(np.random.default_rng(21), T, B, 0.01, reward_scale=3.0)
    tree = _tree(batch, extra=("past_return",))
    fused = FusedBatchTransforms([NormalizeRewards(tree, 0.99), ClipRewards(-2, 2), EstimateAdvantage(0.99, 0.95),
                                  NormalizeAdvantage(tree)], use_cuda_graph=False)
    before = _lib.launch_counters()
    fused(tree)
    after = _lib.launch_counters()
    assert after[1] - before[1] == 1 and after[0] - before[0] == 2, "expected past-return scan + resident kernel"
    r, past = batch["reward"].copy(), np.empty_like(batch["reward"])
    stats = O.RunningMeanStd(())
    O.normalize_rewards(r, batch["done"], past, np.zeros(B, np.float32), np.zeros(B, bool), stats, 0.99)
    O.clip_rewards(r, -2, 2)
    adv, ret = np.empty_like(r), np.empty_like(r)
    O.compute_gae_advantage(r, batch["value"], batch["done"], batch["bootstrap_value"], 0.99, 0.95, adv, ret)
    O.normalize_advantage(adv)
    for k, ref in (("past_return", past), ("reward", r), ("return_", ret), ("advantage", adv)):
        assert_within_tolerance(tree[k].cpu().numpy(), ref, what=f"chain {k}")


def test_default_policy_keeps_one_gpu_on_the_two_kernel_route():
    from parllel_b200 import _lib
    _lib.load().plb_set_tuning(12, 1)
    batch = make_batch(np.random.default_rng(4), 128, 256, 0.01)
    adv_ref, ret_ref = _oracle_step(batch, 0.99, 0.95)
    tree = _tree(batch)
    before = _lib.launch_counters()
    _fused(tree, 0.95)(tree)
    after = _lib.launch_counters()
    assert after[1] == before[1] and after[0] - before[0] == 2
    assert_within_tolerance(tree["advantage"].cpu().numpy(), adv_ref, what="advantage")


def test_batches_too_large_for_tensor_memory_take_the_two_kernel_route():
    from parllel_b200 import _lib
    T, B = 2049, 32   # 17 tiles of 128 rows in one strip: one more than a CTA's tensor memory holds
    batch = make_batch(np.random.default_rng(3), T, B, 0.01)
    adv_ref, ret_ref = _oracle_step(batch, 0.99, 0.95)
    tree = _tree(batch)
    before = _lib.launch_counters()
    _fused(tree, 0.95)(tree)
    assert _lib.launch_counters()[1] == before[1]
    assert_within_tolerance(tree["advantage"].cpu().numpy(), adv_ref, what="advantage")
    assert_within_tolerance(tree["return_"].cpu().numpy(), ret_ref, what="return_")
