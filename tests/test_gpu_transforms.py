"""GPU parity: NormalizeRewards / ClipRewards / NormalizeAdvantage / NormalizeObservations and the
running statistics, against the reference's golden outputs and the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.util import assert_within_tolerance, make_batch

pytestmark = pytest.mark.gpu


def _make_device_tree(T, B, with_valid):
    """Device mirror of the sample tree parllel/patterns.py allocates (done / past_return padded)."""
    from parllel_b200 import ArrayDict, DeviceArray
    tree = ArrayDict({
        "reward": DeviceArray((T, B), np.float32),
        "done": DeviceArray((T, B), np.bool_, padding=1),
        "agent_info": {"value": DeviceArray((T, B), np.float32)},
        "bootstrap_value": DeviceArray((B,), np.float32),
        "advantage": DeviceArray((T, B), np.float32),
        "return_": DeviceArray((T, B), np.float32),
        "past_return": DeviceArray((T, B), np.float32, padding=1),
    })
    if with_valid:
        tree["valid"] = DeviceArray((T, B), np.bool_)
    return tree


@pytest.mark.parametrize("name", ["pipeline_c0", "pipeline_valid", "pipeline_lambda1"])
@pytest.mark.parametrize("leaves", ["device_array", "plain_tensor"])
def test_batch_transform_pipeline_matches_reference(golden, name, leaves):
    """NormalizeRewards -> ClipRewards -> EstimateAdvantage -> NormalizeAdvantage over consecutive
    rotating batches (examples/cartpole/train_ppo.py:101-121), carry rows included."""
    from parllel_b200 import ArrayDict
    from parllel_b200.transforms import ClipRewards, EstimateAdvantage, NormalizeAdvantage, NormalizeRewards
    g = golden(name)
    T, B, p_done, n_batches, with_valid, lam, clip_lo, clip_hi, init_count = g["meta"]
    T, B, with_valid = int(T), int(B), bool(with_valid)
    tree = _make_device_tree(T, B, with_valid)
    if leaves == "plain_tensor":  # no padding anywhere: the transform carries the previous rows itself
        tree = ArrayDict({k: (v if isinstance(v, (torch.Tensor, ArrayDict)) else v.tensor) for k, v in tree.items()})
        tree["agent_info"] = ArrayDict({"value": tree["agent_info"]["value"].tensor})
    transforms = [
        NormalizeRewards(tree, discount=0.99, initial_count=None if init_count < 0 else init_count),
        ClipRewards(reward_min=clip_lo, reward_max=clip_hi),
        EstimateAdvantage(discount=0.99, gae_lambda=lam),
        NormalizeAdvantage(tree),
    ]
    for it in range(int(n_batches)):
        if leaves == "device_array":
            tree.rotate()
        tree["reward"][...] = torch.from_numpy(g[f"b{it}_in_reward"]).cuda()
        tree["agent_info"]["value"][...] = torch.from_numpy(g[f"b{it}_in_value"]).cuda()
        tree["done"][...] = torch.from_numpy(g[f"b{it}_in_done"]).cuda()
        tree["bootstrap_value"][...] = torch.from_numpy(g[f"b{it}_in_bootstrap_value"]).cuda()
        if with_valid:
            tree["valid"][...] = torch.from_numpy(g[f"b{it}_in_valid"]).cuda()
        if leaves == "device_array":  # rotate() must have carried the previous rows into the padding
            assert np.array_equal(tree["past_return"][-1].cpu().numpy(), g[f"b{it}_in_prev_past_return"]) or it > 0
            assert np.array_equal(tree["done"][-1].cpu().numpy(), g[f"b{it}_in_prev_done"])
        for tf in transforms:
            out = tf(tree)
            assert out is tree
        for k in ("past_return", "reward", "return_", "advantage"):
            assert_within_tolerance(np.asarray(tree[k].cpu() if isinstance(tree[k], torch.Tensor) else tree[k]),
                                    g[f"b{it}_out_{k}"], what=f"{name} batch {it} {k}")
        st = transforms[0].return_statistics
        np.testing.assert_allclose([st.mean, st.var, st.count], g[f"b{it}_out_stats"], rtol=2e-6)


@pytest.mark.parametrize("name", ["pipeline_c0", "pipeline_valid", "pipeline_lambda1"])
@pytest.mark.parametrize("leaves", ["device_array", "plain_tensor", "host_numpy"])
@pytest.mark.parametrize("mode", ["eager", "graph", "direct"])
def test_fused_batch_transforms_match_reference(golden, name, leaves, mode):
    """The same chain through FusedBatchTransforms (one C call; rebuilt every call, replayed as a CUDA graph, or
    the cached descriptor launched directly with programmatic dependent launch)."""
    graph = mode != "eager"
    from parllel_b200 import ArrayDict
    from parllel_b200.transforms import (ClipRewards, EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage,
                                         NormalizeRewards)
    g = golden(name)
    T, B, p_done, n_batches, with_valid, lam, clip_lo, clip_hi, init_count = g["meta"]
    T, B, with_valid = int(T), int(B), bool(with_valid)
    tree = _make_device_tree(T, B, with_valid)
    if leaves != "device_array":
        tree = ArrayDict({k: (v if isinstance(v, (torch.Tensor, ArrayDict)) else v.tensor) for k, v in tree.items()})
        tree["agent_info"] = ArrayDict({"value": tree["agent_info"]["value"].tensor})
    if leaves == "host_numpy":
        tree = tree.apply(lambda t: t.cpu().numpy())
    fused = FusedBatchTransforms([
        NormalizeRewards(tree, discount=0.99, initial_count=None if init_count < 0 else init_count),
        ClipRewards(reward_min=clip_lo, reward_max=clip_hi),
        EstimateAdvantage(discount=0.99, gae_lambda=lam),
        NormalizeAdvantage(tree),
    ], use_cuda_graph=graph, launch=mode if graph else None)
    reps = 3 if (graph and leaves == "plain_tensor") else 1  # replays must not depend on stale state
    for it in range(int(n_batches)):
        if leaves == "device_array":
            tree.rotate()
        for rep in range(reps):
            if rep > 0:  # undo the state side effects so the same batch can be replayed through the graph
                fused.norm_rewards.return_statistics.load_state(*saved_stats)
                fused.norm_rewards._carry_return.copy_(saved_carry[0])
                fused.norm_rewards._carry_done.copy_(saved_carry[1])
            elif leaves == "plain_tensor" and fused.norm_rewards._carry_return is not None:
                saved_carry = (fused.norm_rewards._carry_return.clone(), fused.norm_rewards._carry_done.clone())
            if rep == 0:
                st0 = fused.norm_rewards.return_statistics
                saved_stats = (st0.mean.copy(), st0.var.copy(), float(st0.count))
                if leaves == "plain_tensor" and fused.norm_rewards._carry_return is None:
                    saved_carry = (torch.zeros(B, device="cuda"), torch.zeros(B, dtype=torch.bool, device="cuda"))
            for k, src in (("reward", "reward"), ("done", "done"), ("bootstrap_value", "bootstrap_value")):
                val = g[f"b{it}_in_{src}"]
                tree[k][...] = val if leaves == "host_numpy" else torch.from_numpy(val).cuda()
            val = g[f"b{it}_in_value"]
            tree["agent_info"]["value"][...] = val if leaves == "host_numpy" else torch.from_numpy(val).cuda()
            if with_valid:
                val = g[f"b{it}_in_valid"]
                tree["valid"][...] = val if leaves == "host_numpy" else torch.from_numpy(val).cuda()
            assert fused(tree) is tree
        for k in ("past_return", "reward", "return_", "advantage"):
            leaf = tree[k]
            got = leaf.cpu().numpy() if isinstance(leaf, torch.Tensor) else np.asarray(leaf)
            assert_within_tolerance(got, g[f"b{it}_out_{k}"], what=f"{name} batch {it} {k}")
        st = fused.norm_rewards.return_statistics
        np.testing.assert_allclose([st.mean, st.var, st.count], g[f"b{it}_out_stats"], rtol=2e-6)


@pytest.mark.parametrize("name", ["pipeline_c0", "pipeline_lambda1"])
def test_host_batch_pipeline_overlapped_submits(golden, name):
    """Host trees through HostBatchPipeline: batch i+1 is submitted before batch i is waited on."""
    from parllel_b200 import ArrayDict
    from parllel_b200.host_pipeline import HostBatchPipeline
    from parllel_b200.transforms import (ClipRewards, EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage,
                                         NormalizeRewards)
    g = golden(name)
    T, B, p_done, n_batches, with_valid, lam, clip_lo, clip_hi, init_count = g["meta"]
    T, B, n_batches = int(T), int(B), int(n_batches)

    def host_tree(it):
        return ArrayDict({
            "reward": g[f"b{it}_in_reward"].copy(), "done": g[f"b{it}_in_done"].copy(),
            "agent_info": {"value": g[f"b{it}_in_value"].copy()},
            "bootstrap_value": g[f"b{it}_in_bootstrap_value"].copy(),
            "advantage": np.zeros((T, B), np.float32), "return_": np.zeros((T, B), np.float32),
            "past_return": np.zeros((T, B), np.float32)})

    trees = [host_tree(it) for it in range(n_batches)]
    fused = FusedBatchTransforms([
        NormalizeRewards(trees[0], discount=0.99, initial_count=None if init_count < 0 else init_count),
        ClipRewards(reward_min=clip_lo, reward_max=clip_hi),
        EstimateAdvantage(discount=0.99, gae_lambda=lam),
        NormalizeAdvantage(trees[0]),
    ])
    pipe = HostBatchPipeline(fused)
    handles = [pipe.submit(t) for t in trees]  # all in flight at once (more batches than slots)
    for it, h in enumerate(handles):
        out = h.wait()
        assert out is trees[it]
        for k in ("past_return", "reward", "return_", "advantage"):
            assert_within_tolerance(out[k], g[f"b{it}_out_{k}"], what=f"{name} batch {it} {k}")


def test_fused_subsets_and_order_errors():
    from parllel_b200.transforms import ClipRewards, EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage
    with pytest.raises(ValueError):
        FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), ClipRewards(-1, 1)])
    T, B = 200, 48
    batch = make_batch(np.random.default_rng(4), T, B, 0.03)
    dev = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
    def tree():
        return {"reward": dev["reward"].clone(), "done": dev["done"], "agent_info": {"value": dev["value"]},
                "bootstrap_value": dev["bootstrap_value"], "advantage": torch.zeros(T, B, device="cuda"),
                "return_": torch.zeros(T, B, device="cuda")}
    a, b = tree(), tree()
    seq = [ClipRewards(-0.5, None), EstimateAdvantage(0.99, 0.9), NormalizeAdvantage(a)]
    for tf in seq:
        tf(a)
    FusedBatchTransforms([ClipRewards(-0.5, None), EstimateAdvantage(0.99, 0.9), NormalizeAdvantage(b)])(b)
    for k in ("reward", "advantage", "return_"):
        assert_within_tolerance(b[k].cpu().numpy(), a[k].cpu().numpy(), what=k)
    c = tree()
    FusedBatchTransforms([EstimateAdvantage(0.99, 0.9)])(c)
    EstimateAdvantage(0.99, 0.9)(a := tree())
    assert torch.equal(c["advantage"], a["advantage"]) and torch.equal(c["return_"], a["return_"])


@pytest.mark.parametrize("name", ["norm_obs", "norm_obs_valid"])
def test_normalize_observations_matches_reference(golden, name):
    from parllel_b200 import ArrayDict, DeviceArray
    from parllel_b200.transforms import NormalizeObservations
    g = golden(name)
    B, n_steps, with_valid, init_count = g["meta"]
    B, n_steps = int(B), int(n_steps)
    shape = tuple(int(x) for x in g["obs_shape"])
    tree = ArrayDict({"observation": DeviceArray((n_steps, B), np.float32, feature_shape=shape, padding=1)})
    tree["observation"][...] = torch.from_numpy(g["in_observation"]).cuda()
    if with_valid:
        tree["valid"] = DeviceArray((n_steps, B), np.bool_)
        tree["valid"][...] = torch.from_numpy(g["in_valid"]).cuda()
    tf = NormalizeObservations(tree, obs_shape=shape, initial_count=None if init_count < 0 else init_count)
    for t in range(n_steps):
        tf(tree[t])  # step transform on sample_tree[t]
        np.testing.assert_allclose(tf.obs_statistics.mean, g[f"s{t}_mean"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(tf.obs_statistics.var, g[f"s{t}_var"], rtol=2e-6)
        np.testing.assert_allclose(tf.obs_statistics.count, g[f"s{t}_count"], rtol=1e-15)
    got = np.asarray(tree["observation"])
    np.testing.assert_allclose(got, g["out_observation"], rtol=1e-5, atol=1e-5)


def test_normalize_observations_batched_entry_equals_stepwise(golden):
    from parllel_b200.transforms import NormalizeObservations
    g = golden("norm_obs")
    shape = tuple(int(x) for x in g["obs_shape"])
    obs = torch.from_numpy(g["in_observation"]).cuda()
    tf = NormalizeObservations({"observation": obs}, obs_shape=shape, initial_count=10000)
    tf.normalize_batch(obs)
    np.testing.assert_allclose(obs.cpu().numpy(), g["out_observation"], rtol=1e-5, atol=1e-5)


def test_normalize_advantage_multi_agent_matches_golden(golden):
    from parllel_b200.transforms import NormalizeAdvantage
    g = golden("multi_advantage")
    for i in range(int(g["n_cases"])):
        adv = torch.from_numpy(g[f"{i}_advantage"]).cuda()
        tree = {"advantage": adv}
        NormalizeAdvantage(tree)(tree)
        assert_within_tolerance(adv.cpu().numpy(), g[f"{i}_norm_advantage"], what=f"case {i}")


@pytest.mark.parametrize("shape", [(1, 3), (7, 5), (128, 16), (333, 77), (1024, 4096)])
@pytest.mark.parametrize("masked", [False, True])
def test_batch_moments_vs_numpy_float64(shape, masked):
    from parllel_b200 import _lib
    rng = np.random.default_rng(shape[0] + shape[1])
    x = (100.0 + 3.0 * rng.standard_normal(shape)).astype(np.float32)  # large mean: cancellation stress
    valid = rng.random(shape) < 0.6 if masked else None
    if masked:
        valid[0, 0] = True
    lib = _lib.load()
    xd = torch.from_numpy(x).cuda()
    vd = torch.from_numpy(valid).cuda() if masked else None
    out = torch.zeros(3, dtype=torch.float64, device="cuda")
    ws = torch.zeros(lib.plb_batch_moments_workspace_bytes(), dtype=torch.uint8, device="cuda")
    for _ in range(2):  # second call checks the workspace is left reusable
        _lib.check(lib.plb_batch_moments_f32(xd.data_ptr(), shape[1], shape[0], shape[1],
                                             vd.data_ptr() if masked else 0, shape[1], out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    n, mean, m2 = out.cpu().numpy()
    sel = x[valid].astype(np.float64) if masked else x.astype(np.float64).ravel()
    assert n == sel.size
    np.testing.assert_allclose(mean, sel.mean(), rtol=1e-13)
    np.testing.assert_allclose(m2 / n, sel.var(), rtol=1e-11)


@pytest.mark.parametrize("rows,cols", [(1, 4), (5, 3), (64, 192), (1024, 21168), (100, 10)])
def test_column_moments_vs_numpy_float64(rows, cols):
    from parllel_b200 import _lib
    rng = np.random.default_rng(rows + cols)
    x = (5.0 + 3.0 * rng.standard_normal((rows, cols))).astype(np.float32)
    valid = rng.random(rows) < 0.7
    valid[0] = True
    lib = _lib.load()
    xd, vd = torch.from_numpy(x).cuda(), torch.from_numpy(valid).cuda()
    cnt = torch.zeros(1, dtype=torch.float64, device="cuda")
    mean = torch.zeros(cols, dtype=torch.float64, device="cuda")
    m2 = torch.zeros(cols, dtype=torch.float64, device="cuda")
    ws = torch.zeros(lib.plb_column_moments_workspace_bytes(rows, cols), dtype=torch.uint8, device="cuda")
    for use_valid in (False, True):
        _lib.check(lib.plb_column_moments_f32(xd.data_ptr(), cols, rows, cols, vd.data_ptr() if use_valid else 0,
                                              cnt.data_ptr(), mean.data_ptr(), m2.data_ptr(), ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream().cuda_stream))
        sel = x[valid] if use_valid else x
        assert float(cnt) == sel.shape[0]
        np.testing.assert_allclose(mean.cpu().numpy(), sel.astype(np.float64).mean(0), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(m2.cpu().numpy() / sel.shape[0], sel.astype(np.float64).var(0), rtol=1e-10, atol=1e-12)


def test_running_mean_std_reference_unit_test_on_device():
    """The reference's own RunningMeanStd test (transforms/tests/test_running_mean_std.py:40-59),
    a representative subset of its grid, run through the device implementation."""
    from parllel_b200.transforms import RunningMeanStd
    for batch_size, n_batches in [(1, 100), (10, 100), (100, 10), (100, 1)]:
        for mu, sd in [(0., 1.0), (5., 5.), (100., 5.), (100., 200.)]:
            rng = np.random.default_rng(seed=42)
            stats = RunningMeanStd(shape=(1,))
            all_data = np.zeros((batch_size * n_batches, 1), np.float64)
            for i in range(n_batches):
                batch = rng.normal(mu, sd, size=(batch_size, 1))
                stats.update(torch.from_numpy(batch.astype(np.float32)).cuda(), f32_flow=False)
                all_data[i * batch_size:(i + 1) * batch_size] = batch.astype(np.float32)
            assert np.allclose(np.mean(all_data, axis=0), stats.mean)
            assert np.allclose(np.var(all_data, axis=0), stats.var, rtol=1e-3)
            assert np.allclose(all_data.shape[0], stats.count, atol=1e-2)


def test_update_from_moments_bit_exact_vs_oracle():
    """float64 Chan merge: same operation order as running_mean_std.py:23-30 -> identical bits."""
    from parllel_b200 import _lib
    rng = np.random.default_rng(0)
    n = 1000
    lib = _lib.load()
    for f32 in (False, True):
        bm = rng.standard_normal(n) * 5
        bv = rng.random(n) * 9 + 0.1
        if f32:
            bm, bv = bm.astype(np.float32).astype(np.float64), bv.astype(np.float32).astype(np.float64)
        bn = 48.0
        stats = O.RunningMeanStd((n,), initial_count=3.5)
        stats.mean[:] = rng.standard_normal(n)
        stats.var[:] = rng.random(n) + 0.5
        mean0, var0 = stats.mean.copy(), stats.var.copy()
        stats.update_from_moments(bm.astype(np.float32) if f32 else bm, bv.astype(np.float32) if f32 else bv, bn)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float64)).cuda()
        mean, var, cnt = d(mean0), d(var0), d(np.array([3.5]))
        bcnt, bmean, bm2 = d(np.array([bn])), d(bm), d(bv * bn)
        _lib.check(lib.plb_update_from_moments_f64(bcnt.data_ptr(), bmean.data_ptr(), bm2.data_ptr(), mean.data_ptr(),
                                                   var.data_ptr(), cnt.data_ptr(), n,
                                                   _lib.MOMENTS_F32_FLOW if f32 else 0,
                                                   torch.cuda.current_stream().cuda_stream))
        assert np.array_equal(mean.cpu().numpy(), stats.mean)
        if not f32:  # bv*bn/bn is not always exactly bv, so only the float64 flow can be bit-compared on var
            np.testing.assert_allclose(var.cpu().numpy(), stats.var, rtol=1e-15)
        else:
            np.testing.assert_allclose(var.cpu().numpy(), stats.var, rtol=1e-7)
        assert float(cnt) == float(stats.count)


def test_clip_rewards_semantics():
    from parllel_b200.transforms import ClipRewards
    with pytest.raises(ValueError):
        ClipRewards()
    x = np.random.default_rng(0).standard_normal((37, 19)).astype(np.float32) * 4
    for lo, hi in [(-1.0, None), (None, 2.0), (-0.5, 0.25)]:
        t = {"reward": torch.from_numpy(x.copy()).cuda()}
        ClipRewards(lo, hi)(t)
        assert np.array_equal(t["reward"].cpu().numpy(), np.clip(x, lo, hi))


def test_constructor_errors_match_reference():
    from parllel_b200.transforms import NormalizeObservations, NormalizeRewards
    tree = {"reward": torch.zeros(2, 2), "observation": torch.zeros(2, 2, 3)}
    with pytest.raises(ValueError):
        NormalizeRewards(tree, discount=0.99, initial_count=0.5)
    with pytest.raises(ValueError):
        NormalizeObservations(tree, obs_shape=(3,), initial_count=0.0)
    with pytest.raises(NotImplementedError):
        NormalizeObservations({"observation": {"a": 1}}, obs_shape=(3,))
    with pytest.raises(NotImplementedError):
        NormalizeRewards({"reward": {"a": 1}}, discount=0.9)


def test_c2_shape_observation_step_vs_oracle():
    """BASELINE config C2 shape (B=1024, obs 3x84x84), two dependent steps, vs the oracle."""
    from parllel_b200.transforms import NormalizeObservations
    rng = np.random.default_rng(2)
    shape = (3, 84, 84)
    obs = (5.0 + 3.0 * rng.standard_normal((2, 1024) + shape)).astype(np.float32)
    ref = obs.copy()
    stats = O.RunningMeanStd(shape, initial_count=10000)
    for t in range(2):
        O.normalize_observations_step(ref[t], stats)
    d = torch.from_numpy(obs).cuda()
    tf = NormalizeObservations({"observation": d}, obs_shape=shape, initial_count=10000)
    tf.normalize_batch(d)
    np.testing.assert_allclose(d.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(tf.obs_statistics.var, stats.var, rtol=5e-6)


@pytest.mark.parametrize("variant", [1, 2, 4])
@pytest.mark.parametrize("B,F,masked", [(64, 192, False), (300, 52, True), (4100, 64, False), (17, 8, True),
                                        (1024, 200, False), (256, 100, True), (1000, 64, False)])
def test_observation_step_kernel_variants_vs_oracle(variant, B, F, masked):
    """The single-pass designs (shared-memory slab, two-pass L2 slab, two-CTA cluster slab) and the
    fall-backs they defer to (the cluster slab needs an even B >= 128 and F >= 64)."""
    from parllel_b200 import _lib
    from parllel_b200.transforms import NormalizeObservations
    rng = np.random.default_rng(B + F)
    obs = (5.0 + 3.0 * rng.standard_normal((3, B, F))).astype(np.float32)
    valid = (rng.random((3, B)) < 0.7) if masked else None
    if masked:
        valid[:, 0] = True
    ref = obs.copy()
    stats = O.RunningMeanStd((F,), initial_count=50)
    for t in range(3):
        O.normalize_observations_step(ref[t], stats, valid=None if valid is None else valid[t])
    lib = _lib.load()
    lib.plb_set_tuning(3, variant)
    try:
        d = torch.from_numpy(obs).cuda()
        tree = {"observation": d}
        if masked:
            tree["valid"] = torch.from_numpy(valid).cuda()
        tf = NormalizeObservations(tree, obs_shape=(F,), initial_count=50)
        tf.normalize_batch(d, valid=tree.get("valid"))
    finally:
        lib.plb_set_tuning(3, 0)
    # the reference's float32 np.mean/np.var over axis 0 adds rows sequentially, so its own error grows
    # with B; the device statistics are checked tightly against a float64 restatement and the normalised
    # output against the (float32-moment) oracle with the stated 1e-5 tolerance scaled by that error
    exact = O.RunningMeanStd((F,), initial_count=50)
    for t in range(3):
        sel = obs[t][valid[t]] if masked else obs[t]
        exact.update_from_moments(sel.astype(np.float64).mean(0), sel.astype(np.float64).var(0), sel.shape[0])
    np.testing.assert_allclose(tf.obs_statistics.mean, exact.mean, rtol=2e-7, atol=1e-7)
    np.testing.assert_allclose(tf.obs_statistics.var, exact.var, rtol=1e-6)
    tol = 1e-5 * max(1.0, B / 1000.0)
    np.testing.assert_allclose(d.cpu().numpy(), ref, rtol=tol, atol=tol)


def test_observation_clip_extension():
    """clip=c (extension) equals the reference normalisation followed by np.clip, on every kernel variant."""
    from parllel_b200 import _lib
    from parllel_b200.transforms import NormalizeObservations
    rng = np.random.default_rng(0)
    obs = (5.0 + 3.0 * rng.standard_normal((2, 96, 40))).astype(np.float32)
    ref = obs.copy()
    stats = O.RunningMeanStd((40,), initial_count=10)
    for t in range(2):
        O.normalize_observations_step(ref[t], stats)
    ref = np.clip(ref, -1.5, 1.5)
    for variant in (0, 1, 2, 4):
        _lib.load().plb_set_tuning(3, variant)
        d = torch.from_numpy(obs).cuda()
        NormalizeObservations({"observation": d}, obs_shape=(40,), initial_count=10, clip=1.5).normalize_batch(d)
        _lib.load().plb_set_tuning(3, 0)
        np.testing.assert_allclose(d.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
        assert float(d.abs().max()) <= 1.5


class _HostPaddedArray:
    """Stand-in for a parllel ``Array`` on the host (numpy base with leading padding rows, window view
    through ``__array__``, ``arr[-1]`` addressing the padding row before the window, ``rotate()``):
    what the drop-in path duck-types on when the real class is not importable (GPU box)."""

    def __init__(self, T, B, dtype, padding):
        self.padding, self._T = padding, T
        self._base = np.zeros((T + 2 * padding, B), dtype)

    @property
    def shape(self):
        return (self._T,) + self._base.shape[1:]

    batch_shape = shape

    def __array__(self, dtype=None, copy=None):
        return self._base[self.padding: self.padding + self._T]

    def __getitem__(self, i):
        assert isinstance(i, int)
        out = _HostPaddedArray.__new__(_HostPaddedArray)
        out.padding, out._T, out._base = 0, 1, self._base[self.padding + i][None]
        out.__class__ = _HostRow
        return out

    def __setitem__(self, key, value):
        np.asarray(self)[key] = value

    def rotate(self):
        p = self.padding
        if p:
            self._base[0: 2 * p] = self._base[self._T: self._T + 2 * p].copy()


class _HostRow(_HostPaddedArray):
    def __array__(self, dtype=None, copy=None):
        return self._base[0]

    @property
    def shape(self):
        return self._base.shape[1:]


def test_host_array_leaves_with_padding_rows(golden):
    """Drop-in path on host leaves that carry parllel's padding semantics: the carry rows
    (past_return[-1], done[-1]) are read from the padding, results are written back in place."""
    from parllel_b200 import ArrayDict
    from parllel_b200.transforms import ClipRewards, EstimateAdvantage, NormalizeAdvantage, NormalizeRewards
    g = golden("pipeline_c0")
    T, B, p_done, n_batches, with_valid, lam, clip_lo, clip_hi, init_count = g["meta"]
    T, B = int(T), int(B)
    tree = ArrayDict({
        "reward": _HostPaddedArray(T, B, np.float32, 0), "done": _HostPaddedArray(T, B, np.bool_, 1),
        "agent_info": {"value": _HostPaddedArray(T, B, np.float32, 0)},
        "bootstrap_value": np.zeros(B, np.float32),
        "advantage": _HostPaddedArray(T, B, np.float32, 0), "return_": _HostPaddedArray(T, B, np.float32, 0),
        "past_return": _HostPaddedArray(T, B, np.float32, 1)})
    tfs = [NormalizeRewards(tree, discount=0.99), ClipRewards(clip_lo, clip_hi), EstimateAdvantage(0.99, lam),
           NormalizeAdvantage(tree)]
    for it in range(int(n_batches)):
        for k in ("reward", "done", "advantage", "return_", "past_return"):
            tree[k].rotate()
        tree["reward"][...] = g[f"b{it}_in_reward"]
        tree["done"][...] = g[f"b{it}_in_done"]
        tree["agent_info"]["value"][...] = g[f"b{it}_in_value"]
        tree["bootstrap_value"][...] = g[f"b{it}_in_bootstrap_value"]
        for tf in tfs:
            assert tf(tree) is tree
        for k in ("past_return", "reward", "return_", "advantage"):
            assert_within_tolerance(np.asarray(tree[k]), g[f"b{it}_out_{k}"], what=f"batch {it} {k}")


def test_zero_extent_calls_are_noops():
    """Empty inputs: every entry point returns success without touching memory."""
    from parllel_b200 import _lib
    lib = _lib.load()
    ws = torch.zeros(lib.plb_batch_moments_workspace_bytes(), dtype=torch.uint8, device="cuda")
    out = torch.full((3,), -1.0, dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    assert lib.plb_compute_discount_return_f32(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 0, 1, 0.99, 0, s) == 0
    assert lib.plb_compute_past_discount_return_f32(0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 0.99, 0, s) == 0
    assert lib.plb_scale_clip_f32(0, 0, 0, 7, 0, 1e-6, -1.0, 1.0, s) == 0
    assert lib.plb_normalize_advantage_f32(0, 0, 5, 0, 1, out.data_ptr(), 1e-6, s) == 0
    assert lib.plb_normalize_columns_f32(0, 0, 0, 9, 0, 0, 1e-6, s) == 0
    assert lib.plb_gather_tree(None, 0, 0, 0, 0, s) == 0
    assert lib.plb_replay_sample_indices(out.data_ptr(), 8, 2, 3, 0, 0, 0, 0, 1, out.data_ptr(), out.data_ptr(), s) == 0
    # moments of an empty region: count 0 (mean/M2 0), workspace left reusable
    _lib.check(lib.plb_batch_moments_f32(0, 0, 0, 0, 0, 0, out.data_ptr(), ws.data_ptr(), ws.numel(), s))
    assert out.cpu().tolist() == [0.0, 0.0, 0.0]


def test_metrics_vs_reference_golden(golden):
    """explained_variance / RunningMeanStdModel / valid_mask on the device against the reference's outputs."""
    from parllel_b200.metrics import explained_variance, valid_mask
    from parllel_b200.models import RunningMeanStdModel
    g = golden("metrics")
    for i in range(3):
        ev = explained_variance(torch.from_numpy(g[f"ev{i}_pred"]).cuda(), torch.from_numpy(g[f"ev{i}_true"]).cuda())
        ref = float(g[f"ev{i}"])
        assert abs(ev - ref) <= 2e-6 * max(1.0, abs(ref)), (i, ev, ref)
        assert abs(ev - O.explained_variance(g[f"ev{i}_pred"], g[f"ev{i}_true"])) <= 1e-9
    const = torch.full((100,), 3.0, device="cuda")
    assert np.isnan(explained_variance(const * 0.5, const))
    model = RunningMeanStdModel((5, 3))
    for k in range(3):
        model.update(torch.from_numpy(g["rms_x"][k]).cuda())
        np.testing.assert_allclose(model.mean.cpu().numpy(), g[f"rms{k}_mean"], rtol=2e-6)
        np.testing.assert_allclose(model.var.cpu().numpy(), g[f"rms{k}_var"], rtol=1e-5)
        assert float(model.count) == float(g[f"rms{k}_count"])
    got = valid_mask(torch.from_numpy(g["valid_done"]).cuda())
    np.testing.assert_array_equal(got.cpu().numpy(), g["valid_mask"])


@pytest.mark.parametrize("T,B", [(1, 1), (9, 300), (130, 4097)])
def test_valid_mask_vs_oracle(T, B):
    from parllel_b200.metrics import valid_mask
    rng = np.random.default_rng(T + B)
    done = rng.random((T, B)) < 0.05
    base = torch.zeros((T, B + 3), dtype=torch.bool, device="cuda")  # strided output
    out = valid_mask(torch.from_numpy(done).cuda(), out=base[:, :B])
    np.testing.assert_array_equal(out.cpu().numpy(), O.valid_prefix(done))
    assert not bool(base[:, B:].any())


def test_c2_full_size_observation_step_properties():
    """BASELINE config C2 step shape (B=1024, F=3*84*84): two dependent steps; the running statistics and
    the normalised output against a float64 restatement of norm_obs.py:57-76 evaluated with torch on the
    same device (an independent checker at a size the CPU oracle would take minutes for), and the three
    single-pass kernel designs against each other."""
    from parllel_b200 import _lib
    from parllel_b200.transforms import NormalizeObservations
    B, shape = 1024, (3, 84, 84)
    F = int(np.prod(shape))
    g = torch.Generator(device="cuda").manual_seed(5)
    obs = 5.0 + 3.0 * torch.randn((2, B) + shape, device="cuda", generator=g)
    mean = torch.zeros(F, dtype=torch.float64, device="cuda")
    var = torch.ones(F, dtype=torch.float64, device="cuda")
    count = 10000.0
    expect = []
    for t in range(2):
        x = obs[t].reshape(B, F).to(torch.float64)
        bm, bv = x.mean(0), x.var(0, unbiased=False)
        delta, total = bm - mean, count + B
        mean = mean + delta * B / total
        var = (var * count + bv * B + delta ** 2 * count * B / total) / total
        count = total
        expect.append(((x - mean) / torch.sqrt(var + 1e-6)).to(torch.float32))
    results = {}
    for variant in (1, 2, 4):
        _lib.load().plb_set_tuning(3, variant)
        try:
            d = obs.clone()
            tf = NormalizeObservations({"observation": d}, obs_shape=shape, initial_count=10000)
            for t in range(2):
                tf._step(d[t], None, d.device)
        finally:
            _lib.load().plb_set_tuning(3, 0)
        np.testing.assert_allclose(tf.obs_statistics.mean.reshape(-1), mean.cpu().numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(tf.obs_statistics.var.reshape(-1), var.cpu().numpy(), rtol=2e-6)
        assert float(tf.obs_statistics.count) == count
        for t in range(2):
            err = (d[t].reshape(B, F) - expect[t]).abs().max().item()
            assert err <= 1e-5, (variant, t, err)
        results[variant] = d
    assert (results[1] - results[2]).abs().max().item() <= 2e-6
    assert (results[1] - results[4]).abs().max().item() <= 2e-6


@pytest.mark.parametrize("launch", ["graph", "direct"])
def test_strided_bootstrap_value_is_read_afresh_on_every_call(launch):
    """A row vector that has to be COPIED to become dense (here: every second element of a wider buffer) must not be
    frozen into a bound descriptor or a captured graph: every call reads what the tree holds at that moment."""
    from parllel_b200 import ArrayDict
    from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage
    T, B = 64, 128
    rng = np.random.default_rng(21)
    batch = make_batch(rng, T, B, 0.02)
    wide = torch.zeros(2 * B, device="cuda")
    tree = ArrayDict({"reward": torch.from_numpy(batch["reward"]).cuda(), "done": torch.from_numpy(batch["done"]).cuda(),
                      "agent_info": {"value": torch.from_numpy(batch["value"]).cuda()}, "bootstrap_value": wide[::2],
                      "advantage": torch.empty(T, B, device="cuda"), "return_": torch.empty(T, B, device="cuda")})
    fused = FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(tree)], launch=launch)
    for it in range(4):
        boot = rng.standard_normal(B).astype(np.float32)
        wide[::2] = torch.from_numpy(boot).cuda()
        fused(tree)
        adv, ret = np.empty_like(batch["value"]), np.empty_like(batch["value"])
        O.compute_gae_advantage(batch["reward"], batch["value"], batch["done"], boot, 0.99, 0.95, adv, ret)
        O.normalize_advantage(adv)
        assert_within_tolerance(tree["return_"].cpu().numpy(), ret, what=f"call {it} return_")
        assert_within_tolerance(tree["advantage"].cpu().numpy(), adv, what=f"call {it} advantage")


def test_running_mean_std_model_checkpoint_round_trip(golden):
    """restoring a checkpoint must restore the statistics the NEXT update merges into.  A model that is
    saved after two updates and restored into a fresh instance must continue exactly like the original -- and like
    the reference's third update (golden `rms2_*`)."""
    from parllel_b200.models import RunningMeanStdModel
    g = golden("metrics")
    a = RunningMeanStdModel((5, 3))
    for k in range(2):
        a.update(torch.from_numpy(g["rms_x"][k]).cuda())
    saved = {k: v.clone() for k, v in a.state_dict().items()}
    b = RunningMeanStdModel((5, 3))
    b.load_state_dict(saved)
    np.testing.assert_array_equal(b.mean.cpu().numpy(), a.mean.cpu().numpy())
    for m in (a, b):
        m.update(torch.from_numpy(g["rms_x"][2]).cuda())
    assert float(b.count) == float(a.count) == float(g["rms2_count"])
    # the restored working statistics come from the float32 buffers, the original's are float64: equal to float32 rounding
    np.testing.assert_allclose(b.mean.cpu().numpy(), a.mean.cpu().numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(b.var.cpu().numpy(), a.var.cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(b.mean.cpu().numpy(), g["rms2_mean"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(b.var.cpu().numpy(), g["rms2_var"], rtol=1e-5)
