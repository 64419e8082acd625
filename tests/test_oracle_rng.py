"""Pins the oracle's PCG64 + Lemire restatement (oracle/plb_oracle.c) bit-for-bit against
numpy.random.Generator.integers -- the third-party arithmetic behind
parllel/replays/replay.py:44,64,68,70 (numpy is not vendored in the reference)."""
import numpy as np
import pytest

from oracle import oracle as O


@pytest.mark.parametrize("seed", [0, 1, 12345])
def test_integers_stream_matches_numpy(seed):
    gen = np.random.default_rng(seed)
    mine = O.Pcg64.from_numpy(gen)
    # odd sizes exercise the half-word buffer that persists across calls
    for high, n in [(62463, 7), (16, 5), (3, 1001), (1, 4), (2**31 + 11, 33), (62463, 16384),
                    (4294967295, 9), (1000003, 257), (2**32 - 1, 3), (17, 1)]:
        want = gen.integers(0, high, size=(n,))
        got = mine.integers(high, n)
        assert want.dtype == np.int64
        assert np.array_equal(want, got), (high, n)


def test_rejection_path_is_exercised():
    """high just above 2^31 rejects ~half of all draws; stream must still match."""
    gen = np.random.default_rng(99)
    mine = O.Pcg64.from_numpy(gen)
    high = 2**31 + 1
    assert np.array_equal(gen.integers(0, high, size=(5000,)), mine.integers(high, 5000))
    # and the streams stay aligned afterwards
    assert np.array_equal(gen.integers(0, 10, size=(11,)), mine.integers(10, 11))


def test_replay_index_math_matches_numpy_formula():
    size_T, B, cursor, newest, oldest, n = 62464, 16, 384, 0, 1, 4096
    gen = np.random.default_rng(0)
    mine = O.Pcg64.from_numpy(gen)
    t_ref = (gen.integers(0, size_T - oldest - newest, size=(n,)) + cursor + oldest) % size_T
    b_ref = gen.integers(0, B, size=(n,))
    t, b = O.replay_sample_indices(mine, size_T, B, cursor, True, newest, oldest, n)
    assert np.array_equal(t, t_ref) and np.array_equal(b, b_ref)
