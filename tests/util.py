"""Shared helpers for the parity tests."""
import numpy as np

REL_TOL = 1e-5  # BASELINE.json north_star: float32 tolerance 1e-5 relative


def assert_within_tolerance(got, ref, rel=REL_TOL, what=""):
    """|got - ref| <= rel * max(|ref|, rms(ref)) elementwise.

    Pure elementwise rtol is unattainable on near-zero elements even for exact math
    (hard part 1), so the relative tolerance has the RMS of the reference
    tensor as its absolute floor: equivalent to allclose(rtol=rel, atol=rel*rms)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    if ref.size == 0:
        return
    rms = float(np.sqrt(np.mean(ref * ref)))
    bound = rel * np.maximum(np.abs(ref), rms)
    err = np.abs(got - ref)
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} / {bad.size} elements out of tolerance; worst at {i}: "
                             f"got {got[i]!r} ref {ref[i]!r} err {err[i]:.3e} bound {bound[i]:.3e} (rms {rms:.3e})")


def make_batch(rng, T, B, p_done=0.01, inner=None, reward_scale=1.0):
    feat = () if inner is None else (inner,)
    return dict(
        reward=(rng.standard_normal((T, B)) * reward_scale).astype(np.float32),
        value=rng.standard_normal((T, B) + feat).astype(np.float32),
        done=rng.random((T, B)) < p_done,
        bootstrap_value=rng.standard_normal((B,) + feat).astype(np.float32),
    )
