"""GPU parity: EstimateAdvantage / past-return scans vs the golden reference outputs and the oracle.

Bars (BASELINE.json north_star): GAE / returns within 1e-5 relative (floor = rms, tests/util.py);
the SERIAL algorithm reproduces the reference's per-step float32 rounding and must be
BIT-IDENTICAL to the strict-IEEE oracle (oracle/plb_oracle.c)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.util import assert_within_tolerance, make_batch

pytestmark = pytest.mark.gpu


def _device_tree(batch, dev="cuda"):
    from parllel_b200 import ArrayDict
    t = {k: torch.from_numpy(v).to(dev) for k, v in batch.items()}
    return ArrayDict({
        "reward": t["reward"], "done": t["done"], "agent_info": {"value": t["value"]},
        "bootstrap_value": t["bootstrap_value"],
        "advantage": torch.full_like(t["value"], float("nan")),
        "return_": torch.full_like(t["value"], float("nan")),
    })


def _oracle_adv(batch, gamma, lam):
    adv, ret = np.empty_like(batch["value"]), np.empty_like(batch["value"])
    fn = O.compute_discount_return if lam == 1.0 else O.compute_gae_advantage
    fn(batch["reward"], batch["value"], batch["done"], batch["bootstrap_value"], gamma, lam, adv, ret)
    return adv, ret


@pytest.mark.parametrize("algo", ["serial", "chunked", "auto"])
def test_advantage_matches_reference_golden(golden, algo):
    from parllel_b200.transforms import EstimateAdvantage
    g = golden("advantage")
    for i in range(int(g["n_cases"])):
        T, B, p, lam, gamma = g[f"{i}_meta"]
        batch = {k: g[f"{i}_{k}"] for k in ("reward", "value", "done", "bootstrap_value")}
        if algo == "chunked" and (int(B) % 16 != 0):
            continue  # TMA needs 16-byte row strides; auto falls back to serial for these
        tree = _device_tree(batch)
        out = EstimateAdvantage(discount=gamma, gae_lambda=lam, algo=algo)(tree)
        assert out is tree
        assert_within_tolerance(tree["advantage"].cpu().numpy(), g[f"{i}_advantage"], what=f"adv case {i} {algo}")
        assert_within_tolerance(tree["return_"].cpu().numpy(), g[f"{i}_return_"], what=f"ret case {i} {algo}")


@pytest.mark.parametrize("lam", [0.95, 1.0])
@pytest.mark.parametrize("T,B,p", [(1, 16, 0.1), (5, 48, 0.5), (63, 32, 0.0), (64, 64, 1.0), (65, 80, 0.02),
                                   (200, 1000, 0.01), (1024, 512, 0.01), (130, 33, 0.05), (7, 1, 0.3)])
def test_serial_is_bit_identical_to_oracle(T, B, p, lam):
    from parllel_b200.transforms import EstimateAdvantage
    batch = make_batch(np.random.default_rng(T * 1000 + B), T, B, p)
    tree = _device_tree(batch)
    EstimateAdvantage(0.99, lam, algo="serial")(tree)
    adv, ret = _oracle_adv(batch, 0.99, lam)
    assert np.array_equal(tree["advantage"].cpu().numpy(), adv)
    assert np.array_equal(tree["return_"].cpu().numpy(), ret)


@pytest.mark.parametrize("lam", [0.95, 1.0])
@pytest.mark.parametrize("T,B,p", [(1, 16, 0.1), (5, 48, 0.5), (63, 32, 0.0), (64, 64, 1.0), (65, 80, 0.02),
                                   (200, 1008, 0.01), (1024, 512, 0.01), (129, 4096, 0.01), (3000, 16, 0.001)])
def test_chunked_within_tolerance_of_oracle(T, B, p, lam):
    from parllel_b200.transforms import EstimateAdvantage
    batch = make_batch(np.random.default_rng(T * 1000 + B + 1), T, B, p)
    tree = _device_tree(batch)
    EstimateAdvantage(0.99, lam, algo="chunked")(tree)
    adv, ret = _oracle_adv(batch, 0.99, lam)
    assert_within_tolerance(tree["advantage"].cpu().numpy(), adv, what="advantage")
    assert_within_tolerance(tree["return_"].cpu().numpy(), ret, what="return_")


def test_strided_views_into_padded_bases():
    """done is a window into a [T+2, B] padded base (patterns.py:191-198): row stride == B but an
    offset base pointer; value/advantage as column slices of wider arrays (ld > B)."""
    from parllel_b200 import ArrayDict, DeviceArray
    from parllel_b200.transforms import EstimateAdvantage
    T, B = 96, 64
    batch = make_batch(np.random.default_rng(5), T, B, 0.05)
    done = DeviceArray((T, B), np.bool_, padding=1)
    done[...] = batch["done"]
    wide_v = torch.zeros((T, 2 * B), device="cuda")
    wide_v[:, B:] = torch.from_numpy(batch["value"]).cuda()
    wide_a = torch.zeros((T, 3 * B), device="cuda")
    for algo in ("serial", "chunked"):
        tree = ArrayDict({
            "reward": torch.from_numpy(batch["reward"]).cuda(), "done": done,
            "agent_info": {"value": wide_v[:, B:]}, "bootstrap_value": torch.from_numpy(batch["bootstrap_value"]).cuda(),
            "advantage": wide_a[:, B:2 * B], "return_": torch.zeros((T, B), device="cuda"),
        })
        EstimateAdvantage(0.99, 0.95, algo=algo)(tree)
        adv, ret = _oracle_adv(batch, 0.99, 0.95)
        assert_within_tolerance(wide_a[:, B:2 * B].cpu().numpy(), adv, what=algo)
        assert_within_tolerance(tree["return_"].cpu().numpy(), ret, what=algo)
        assert float(wide_a[:, :B].abs().max()) == 0.0 and float(wide_a[:, 2 * B:].abs().max()) == 0.0


def test_multi_agent_trailing_axis_matches_golden(golden):
    from parllel_b200.transforms import EstimateAdvantage
    g = golden("multi_advantage")
    for i in range(int(g["n_cases"])):
        T, B, N, lam, gamma = g[f"{i}_meta"]
        batch = {k: g[f"{i}_{k}"] for k in ("reward", "value", "done", "bootstrap_value")}
        tree = _device_tree(batch)
        EstimateAdvantage(gamma, lam)(tree)
        assert_within_tolerance(tree["advantage"].cpu().numpy(), g[f"{i}_advantage"], what=f"case {i}")
        assert_within_tolerance(tree["return_"].cpu().numpy(), g[f"{i}_return_"], what=f"case {i}")


def test_host_leaves_are_staged_and_written_back():
    """Drop-in path: numpy leaves in, CUDA kernels compute, results land in the caller's arrays."""
    from parllel_b200 import ArrayDict
    from parllel_b200.transforms import EstimateAdvantage
    batch = make_batch(np.random.default_rng(3), 128, 16, 0.01)
    tree = ArrayDict({"reward": batch["reward"], "done": batch["done"], "agent_info": {"value": batch["value"]},
                      "bootstrap_value": batch["bootstrap_value"],
                      "advantage": np.zeros_like(batch["value"]), "return_": np.zeros_like(batch["value"])})
    EstimateAdvantage(0.99, 0.95, algo="serial")(tree)
    adv, ret = _oracle_adv(batch, 0.99, 0.95)
    assert np.array_equal(tree["advantage"], adv) and np.array_equal(tree["return_"], ret)


@pytest.mark.parametrize("algo", ["serial", "chunked"])
@pytest.mark.parametrize("T,B,p", [(1, 16, 0.5), (64, 64, 0.1), (65, 48, 0.02), (500, 1024, 0.01), (77, 5, 0.1)])
def test_past_discount_return(T, B, p, algo):
    from parllel_b200 import _lib
    if algo == "chunked" and B % 16:
        pytest.skip("TMA needs 16-byte row strides")
    rng = np.random.default_rng(T + B)
    reward = rng.standard_normal((T, B)).astype(np.float32)
    done = rng.random((T, B)) < p
    prev_r = rng.standard_normal(B).astype(np.float32)
    prev_d = rng.random(B) < 0.3
    ref = np.empty_like(reward)
    O.compute_past_discount_return(reward, done, prev_r, prev_d, 0.99, ref)
    r, d = torch.from_numpy(reward).cuda(), torch.from_numpy(done).cuda()
    pr, pd = torch.from_numpy(prev_r).cuda(), torch.from_numpy(prev_d).cuda()
    out = torch.full((T, B), float("nan"), device="cuda")
    lib = _lib.load()
    rc = lib.plb_compute_past_discount_return_f32(r.data_ptr(), B, d.data_ptr(), B, pr.data_ptr(), pd.data_ptr(),
                                                  out.data_ptr(), B, T, B, 0.99,
                                                  _lib.ALGO_SERIAL if algo == "serial" else _lib.ALGO_CHUNKED,
                                                  torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    if algo == "serial":
        assert np.array_equal(out.cpu().numpy(), ref)
    else:
        assert_within_tolerance(out.cpu().numpy(), ref)


def test_c1_full_size_properties():
    """BASELINE config C1 (T=1024, B=4096): serial == oracle bit for bit; chunked within tolerance;
    size-independent properties: an all-done mask decouples every step, linearity in reward."""
    from parllel_b200.transforms import EstimateAdvantage
    T, B = 1024, 4096
    batch = make_batch(np.random.default_rng(1), T, B, 0.01)
    adv_ref, ret_ref = _oracle_adv(batch, 0.99, 0.95)
    tree = _device_tree(batch)
    EstimateAdvantage(0.99, 0.95, algo="serial")(tree)
    assert np.array_equal(tree["advantage"].cpu().numpy(), adv_ref)
    tree2 = _device_tree(batch)
    EstimateAdvantage(0.99, 0.95, algo="chunked")(tree2)
    assert_within_tolerance(tree2["advantage"].cpu().numpy(), adv_ref, what="C1 advantage")
    assert_within_tolerance(tree2["return_"].cpu().numpy(), ret_ref, what="C1 return_")
    # all done: advantage[t] = reward[t] - value[t] exactly
    b3 = dict(batch, done=np.ones((T, B), bool))
    tree3 = _device_tree(b3)
    EstimateAdvantage(0.99, 0.95, algo="chunked")(tree3)
    assert np.array_equal(tree3["advantage"].cpu().numpy(), batch["reward"] - batch["value"])
    # linearity: scaling reward, value and bootstrap by 2 scales the advantage by exactly 2
    b4 = dict(batch, reward=batch["reward"] * 2, value=batch["value"] * 2, bootstrap_value=batch["bootstrap_value"] * 2)
    tree4 = _device_tree(b4)
    EstimateAdvantage(0.99, 0.95, algo="chunked")(tree4)
    assert np.array_equal(tree4["advantage"].cpu().numpy(), 2 * tree2["advantage"].cpu().numpy())


def test_c4_full_size_properties():
    """BASELINE config C4 (T=256, B=65536): the bit-exact serial walk (pinned to the oracle on a column
    sample) against the chunked scan; shard invariance (columns are independent, so any contiguous B split
    reproduces the same bits, which is what the multi-GPU path relies on); the fused chain's advantage equals
    EstimateAdvantage + NormalizeAdvantage issued separately."""
    from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage
    T, B = 256, 65536
    batch = make_batch(np.random.default_rng(2), T, B, 0.01)
    tree = _device_tree(batch)
    EstimateAdvantage(0.99, 0.95, algo="serial")(tree)
    adv_serial = tree["advantage"].cpu().numpy()
    cols = np.random.default_rng(3).choice(B, 512, replace=False)
    sub = {k: (np.ascontiguousarray(v[..., cols]) if v.ndim >= 1 else v) for k, v in batch.items()}
    adv_ref, _ = _oracle_adv(sub, 0.99, 0.95)
    assert np.array_equal(adv_serial[:, cols], adv_ref)
    tree2 = _device_tree(batch)
    EstimateAdvantage(0.99, 0.95, algo="chunked")(tree2)
    adv_chunked = tree2["advantage"].cpu().numpy()
    assert_within_tolerance(adv_chunked, adv_serial, what="C4 chunked vs serial")
    # shard invariance: 8 contiguous column blocks scanned separately reproduce the whole-batch bits on the
    # serial walk; the chunked scan may pick another tile height for the narrower shard (float64 carries
    # composed in a different order), so it is held to the stated tolerance
    for r in range(8):
        lo, hi = r * B // 8, (r + 1) * B // 8
        shard = {k: (np.ascontiguousarray(v[..., lo:hi]) if v.ndim >= 1 else v) for k, v in batch.items()}
        t = _device_tree(shard)
        EstimateAdvantage(0.99, 0.95, algo="serial")(t)
        assert np.array_equal(t["advantage"].cpu().numpy(), adv_serial[:, lo:hi]), f"shard {r}"
        t = _device_tree(shard)
        EstimateAdvantage(0.99, 0.95, algo="chunked")(t)
        assert_within_tolerance(t["advantage"].cpu().numpy(), adv_serial[:, lo:hi], what=f"chunked shard {r}")
    # fused chain == separate transforms (same kernels' arithmetic, statistics within float64 rounding)
    tree3, tree4 = _device_tree(batch), _device_tree(batch)
    FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(tree3)])(tree3)
    EstimateAdvantage(0.99, 0.95, algo="chunked")(tree4)
    NormalizeAdvantage(tree4)(tree4)
    a3, a4 = tree3["advantage"].cpu().numpy(), tree4["advantage"].cpu().numpy()
    np.testing.assert_allclose(a3, a4, rtol=2e-6, atol=2e-6)
    assert abs(float(a3.mean())) < 1e-4 and abs(float(a3.std()) - 1.0) < 1e-4
    assert np.array_equal(tree3["return_"].cpu().numpy(), tree4["return_"].cpu().numpy())


def test_empty_batch_is_a_noop():
    from parllel_b200 import _lib
    lib = _lib.load()
    rc = lib.plb_compute_gae_advantage_f32(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 16, 1, 0.99, 0.95, 0, 0)
    assert rc == 0


def test_bad_arguments_fail_loudly():
    from parllel_b200 import _lib
    lib = _lib.load()
    x = torch.zeros((4, 8), device="cuda")
    rc = lib.plb_compute_gae_advantage_f32(x.data_ptr(), 4, x.data_ptr(), 8, x.data_ptr(), 8, x.data_ptr(),
                                           x.data_ptr(), 8, x.data_ptr(), 8, 4, 8, 1, 0.99, 0.95, 0, 0)
    assert rc == -1 and b"stride" in lib.plb_last_error()
    with pytest.raises(_lib.PlbError):
        _lib.check(rc, "gae")


def test_estimate_multi_agent_advantage_problem_types(golden):
    """EstimateMultiAgentAdvantage: independent critics and single critic against the reference's golden
    outputs; markov game (per-agent rewards) against the oracle (the reference's own branch does not run
    on numpy >= 1.24, see parllel_b200/transforms/multi_advantage.py)."""
    from parllel_b200 import ArrayDict
    from parllel_b200.transforms import EstimateMultiAgentAdvantage
    g = golden("multi_advantage")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for i in range(int(g["n_cases"])):  # independent critics
        T, B, N, lam, gamma = g[f"{i}_meta"]
        tree = ArrayDict({"reward": d(g[f"{i}_reward"]), "done": d(g[f"{i}_done"]),
                          "agent_info": {"value": d(g[f"{i}_value"])}, "bootstrap_value": d(g[f"{i}_bootstrap_value"]),
                          "advantage": torch.zeros_like(d(g[f"{i}_value"])), "return_": torch.zeros_like(d(g[f"{i}_value"]))})
        EstimateMultiAgentAdvantage(tree, gamma, lam)(tree)
        assert_within_tolerance(tree["advantage"].cpu().numpy(), g[f"{i}_advantage"], what=f"independent {i}")
        assert_within_tolerance(tree["return_"].cpu().numpy(), g[f"{i}_return_"], what=f"independent {i}")
    # single critic: advantage has a trailing singleton axis
    v = d(g["single_value"])
    tree = ArrayDict({"reward": d(g["single_reward"]), "done": d(g["single_done"]), "agent_info": {"value": v},
                      "bootstrap_value": d(g["single_bootstrap_value"]),
                      "advantage": torch.zeros(v.shape + (1,), device="cuda"), "return_": torch.zeros_like(v)})
    EstimateMultiAgentAdvantage(tree, 0.99, 0.95)(tree)
    assert_within_tolerance(tree["advantage"].cpu().numpy(), g["single_advantage"], what="single critic")
    assert_within_tolerance(tree["return_"].cpu().numpy(), g["single_return_"], what="single critic")
    with pytest.raises(ValueError):
        EstimateMultiAgentAdvantage({"reward": v, "agent_info": {"value": v}, "advantage": v}, 0.99, 0.95)
    # markov game vs the oracle
    T, B, N = 20, 6, 3
    rng = np.random.default_rng(9)
    names = ["zeta", "alpha", "mid"]
    rewards = {n: rng.standard_normal((T, B)).astype(np.float32) for n in names}
    value = rng.standard_normal((T, B, N)).astype(np.float32)
    done = rng.random((T, B)) < 0.1
    boot = rng.standard_normal((B, N)).astype(np.float32)
    tree = ArrayDict({"reward": {n: d(rewards[n]) for n in reversed(names)}, "action": {n: d(rewards[n]) for n in names},
                      "done": d(done), "agent_info": {"value": d(value)}, "bootstrap_value": d(boot),
                      "advantage": torch.zeros(T, B, N, device="cuda"), "return_": torch.zeros(T, B, N, device="cuda")})
    EstimateMultiAgentAdvantage(tree, 0.99, 0.9)(tree)
    stacked = np.stack([rewards[n] for n in names], axis=-1)
    adv, ret = np.empty_like(value), np.empty_like(value)
    O.compute_gae_advantage(stacked.reshape(T, B * N), value.reshape(T, B * N), np.repeat(done, N, axis=1),
                            boot.reshape(-1), 0.99, 0.9, adv.reshape(T, B * N), ret.reshape(T, B * N))
    assert_within_tolerance(tree["advantage"].cpu().numpy(), adv, what="markov game")
    assert_within_tolerance(tree["return_"].cpu().numpy(), ret, what="markov game")
