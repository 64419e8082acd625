"""CPU-only tests: the C-ABI library loads and exports every declared symbol, and the host-side
mirrors of the reference interface (ArrayDict, DeviceArray layout, ReplayBuffer index math)
behave like the reference.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle.ref_loader import load_reference, reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "parllel_b200.h")).read()
    return sorted(set(re.findall(r"PLB_API[^;(]*?\b(plb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from parllel_b200 import _lib
    assert os.path.exists(_lib.library_path()), "libparllel_b200.so not built (python -m parllel_b200.build)"
    lib = ctypes.CDLL(_lib.library_path())
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/parllel_b200.h but not exported"
    bound = _lib.load()
    assert bound.plb_version() == 200


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from parllel_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PlbError, match="no CPU fallback"):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "parllel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libplb_oracle" not in text, f
                assert not re.search(r'#include\s+"[^"]*oracle', text), f


# ----------------------------------------------------------------------------- ArrayDict
def test_array_dict_indexing_and_fanout():
    from parllel_b200 import ArrayDict
    a, b = np.arange(12.).reshape(3, 4), np.arange(24.).reshape(3, 4, 2)
    tree = ArrayDict({"a": a, "nested": {"b": b}})
    assert isinstance(tree["nested"], ArrayDict)
    sub = tree[1]
    assert np.array_equal(sub["a"], a[1]) and np.array_equal(sub["nested"]["b"], b[1])
    sub2 = tree[np.array([0, 2]), np.array([1, 3])]
    assert np.array_equal(sub2["a"], a[[0, 2], [1, 3]]) and sub2["nested"]["b"].shape == (2, 2)
    tree[0] = 7.0
    assert (a[0] == 7).all() and (b[0] == 7).all()
    tree[1] = {"a": np.zeros(4)}
    assert (a[1] == 0).all() and (b[1] != 0).any()
    shapes = tree.shape
    assert shapes["a"] == (3, 4) and shapes["nested"]["b"] == (3, 4, 2)
    copied = tree.copy()  # attribute fan-out + call
    assert isinstance(copied, ArrayDict) and copied["a"] is not a and np.array_equal(copied["a"], a)
    with pytest.raises(IndexError, match="field 'a'"):
        tree[5]
    with pytest.raises(IndexError):
        del tree[0]
    del tree["a"]
    assert list(tree) == ["nested"] and len(tree) == 1
    assert tree.apply(lambda x: x + 1)["nested"]["b"][2, 3, 1] == b[2, 3, 1] + 1
    assert isinstance(ArrayDict({"t": torch.ones(2)}).to_ndarray()["t"], np.ndarray)


# ----------------------------------------------------------------------------- DeviceArray
@pytest.mark.parametrize("padding,full_mult", [(0, 1), (1, 1), (2, 1), (1, 3), (2, 2)])
def test_device_array_layout_semantics(padding, full_mult):
    """Window / padding / rotate contract of parllel/arrays/array.py, checked against a numpy mirror
    (and against the reference's own Array when it is mounted)."""
    from parllel_b200 import DeviceArray
    T, B = 4, 3
    full = T * full_mult
    arr = DeviceArray((T, B), np.float32, padding=padding, full_size=full, device="cpu")
    assert arr.shape == (T, B) and arr.batch_shape == (T, B) and arr.base.shape == (full + 2 * padding, B)
    ref = None
    if reference_available():
        load_reference()
        from parllel import Array
        ref = Array(batch_shape=(T, B), dtype=np.float32, padding=padding, full_size=full)
    counter = 0.0
    for it in range(2 * full_mult + 1):
        vals = np.arange(T * B, dtype=np.float32).reshape(T, B) + counter
        counter += 100
        arr[...] = vals
        assert np.array_equal(np.asarray(arr), vals)
        if padding:
            arr[T] = vals[-1] + 0.5  # trailing padding row (e.g. next observation)
            assert np.array_equal(arr[T].numpy(), vals[-1] + 0.5)
        if ref is not None:
            ref[...] = vals
            if padding:
                ref[T] = vals[-1] + 0.5
            assert np.array_equal(np.asarray(ref), np.asarray(arr))
        arr.rotate()
        if ref is not None:
            ref.rotate()
            assert np.array_equal(ref._base_array, arr.base.numpy()), f"iteration {it}"
            assert ref._current_location[0].start == arr._start
        if padding:
            # after rotate(), the row before the window holds the last row of the previous batch
            assert np.array_equal(arr[-1].numpy(), vals[-1])
            assert np.array_equal(arr[0].numpy(), vals[-1] + 0.5)
    assert arr.full.shape == (full, B)
    with pytest.raises(ValueError):
        arr.full.rotate()
    with pytest.raises(IndexError):
        arr[-padding - 1 - arr._start + padding - 1]
    with pytest.raises(ValueError):
        DeviceArray((4, 2), np.float32, padding=5, device="cpu")
    with pytest.raises(ValueError):
        DeviceArray((4, 2), np.float32, full_size=6, device="cpu")


@pytest.mark.skipif(not reference_available(), reason="reference not mounted")
def test_device_array_from_host_mirrors_reference_array():
    from parllel_b200 import DeviceArray
    load_reference()
    from parllel import Array
    a = Array(batch_shape=(4, 3), dtype=np.float32, feature_shape=(2,), padding=1, full_size=12)
    a.rotate()
    a[...] = np.random.default_rng(0).standard_normal((4, 3, 2)).astype(np.float32)
    d = DeviceArray.from_host(a, device="cpu")
    assert np.array_equal(np.asarray(d), np.asarray(a))
    assert np.array_equal(np.asarray(d.full), np.asarray(a.full))
    assert np.array_equal(d[-1].numpy(), np.asarray(a[-1]))


# ----------------------------------------------------------------------------- ReplayBuffer
def test_replay_index_math_matches_reference_golden(golden):
    from parllel_b200 import ArrayDict, BatchSpec
    from parllel_b200.replays import ReplayBuffer
    g = golden("replay")
    T, B, size_T, _, _ = (int(x) for x in g["meta"])
    tree = ArrayDict({"reward": torch.from_numpy(g["ring_reward"])})
    for name in g["scenarios"]:
        batch, newest, oldest, n_iter, seed, n_draws, cursor, full = (int(x) for x in g[f"{name}_meta"])
        rb = ReplayBuffer(tree, BatchSpec(T, B), size_T=size_T, replay_batch_size=batch,
                          newest_n_samples_invalid=newest, oldest_n_samples_invalid=oldest, device="cpu")
        rb.seed(seed)
        for _ in range(n_iter):
            rb.next_iteration()
        assert (rb._cursor, int(rb._full)) == (cursor, full)
        for d in range(n_draws):
            t_idx, b_idx = rb.sample_indices()
            assert t_idx.dtype == np.int64
            assert np.array_equal(t_idx, g[f"{name}_d{d}_t_index"]) and np.array_equal(b_idx, g[f"{name}_d{d}_b_index"])


def test_transform_constructor_contracts_cpu():
    from parllel_b200.transforms import ClipRewards, EstimateAdvantage, NormalizeObservations, NormalizeRewards
    with pytest.raises(ValueError):
        ClipRewards()
    with pytest.raises(ValueError):
        NormalizeRewards({"reward": np.zeros((2, 2))}, discount=0.99, initial_count=0.5, device="cpu")
    with pytest.raises(ValueError):
        NormalizeObservations({"observation": np.zeros((2, 3))}, obs_shape=(3,), initial_count=0.1, device="cpu")
    with pytest.raises(NotImplementedError):
        NormalizeObservations({"observation": {"k": 1}}, obs_shape=(3,), device="cpu")
    tf = NormalizeRewards({"reward": np.zeros((2, 2)), "valid": 1}, discount=0.99, initial_count=2.0, device="cpu")
    assert tf.only_valid and float(tf.return_statistics.count) == 2.0 and float(tf.return_statistics.var) == 1.0
    ea = EstimateAdvantage(0.99, 1.0)
    assert ea.discount == 0.99 and ea.gae_lambda == 1.0


def test_fused_batch_transforms_launch_modes_and_tuning_keys():
    """The launch path is a constructor choice ("direct": bound descriptor + programmatic dependent launch, the
    default; "graph": CUDA-graph replay; use_cuda_graph=False: rebuilt every call); the experiment switches that
    select kernel variants are plain host state and must be settable without a GPU."""
    from parllel_b200 import _lib
    from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms
    assert FusedBatchTransforms([EstimateAdvantage(0.99, 0.95)]).launch in ("direct", "graph")
    assert FusedBatchTransforms([EstimateAdvantage(0.99, 0.95)], launch="graph").launch == "graph"
    assert FusedBatchTransforms([EstimateAdvantage(0.99, 0.95)], launch="direct").launch == "direct"
    assert FusedBatchTransforms([EstimateAdvantage(0.99, 0.95)], use_cuda_graph=False).launch == "eager"
    with pytest.raises(ValueError):
        FusedBatchTransforms([EstimateAdvantage(0.99, 0.95)], launch="bogus")
    lib = _lib.load()
    for key, value in ((18, 1), (20, 256), (21, 0), (17, 0)):
        assert lib.plb_set_tuning(key, value) == 0
    assert lib.plb_set_tuning(9999, 0) != 0
    assert b"unknown tuning key" in lib.plb_last_error()


def test_valid_mean_matches_the_reference_including_autograd():
    """parllel/torch/utils.py:42-62 -- all three call forms, values and gradients."""
    from parllel_b200.metrics import valid_mean
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 5, 4, generator=g)
    x[1, 2, 0] = 0.0  # a valid element that is exactly zero: the reference's count_nonzero does not count it
    valid = torch.rand(6, 5, generator=g) < 0.7
    valid[1, 2] = True
    cases = [dict(), dict(dim=(0, 1)), dict(valid=valid), dict(valid=valid, dim=(0, 1)), dict(valid=valid, dim=0)]
    # independent expectations
    assert torch.allclose(valid_mean(x), x.mean())
    assert torch.allclose(valid_mean(x, valid), x[valid].mean())
    m = valid[..., None].expand_as(x)
    expect = (x * m).sum(dim=(0, 1)) / ((x * m) != 0).sum(dim=(0, 1))
    assert torch.allclose(valid_mean(x, valid, dim=(0, 1)), expect)
    if not reference_available():
        return
    load_reference()
    from parllel.torch.utils import valid_mean as ref_valid_mean
    for kw in cases:
        a = x.clone().requires_grad_(True)
        b = x.clone().requires_grad_(True)
        got, ref = valid_mean(a, **kw), ref_valid_mean(b, **kw)
        assert torch.equal(got, ref), kw
        got.sum().backward()
        ref.sum().backward()
        assert torch.equal(a.grad, b.grad), kw
