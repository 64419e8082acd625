"""GPU parity: ReplayBuffer -- sampled indices and gathered minibatches are BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

LEAVES = ("observation", "next_observation", "action", "reward", "terminated")


def _ring_tree(g, device):
    from parllel_b200 import ArrayDict
    T, B, size_T, obs_dim, act_dim = (int(x) for x in g["meta"])
    tree = ArrayDict({k: torch.from_numpy(g[f"ring_{k}"]) for k in LEAVES})
    tt, bb = np.meshgrid(np.arange(size_T), np.arange(B), indexing="ij")
    tree["t_index"] = torch.from_numpy(tt.copy())
    tree["b_index"] = torch.from_numpy(bb.copy())
    if device == "cuda":
        tree = tree.apply(lambda t: t.cuda())
    return tree, T, B, size_T


@pytest.mark.parametrize("residency", ["cuda", "host_mirror"])
def test_sample_batch_matches_reference_golden(golden, residency):
    from parllel_b200 import BatchSpec
    from parllel_b200.replays import ReplayBuffer
    g = golden("replay")
    tree, T, B, size_T = _ring_tree(g, "cuda" if residency == "cuda" else "cpu")
    for name in g["scenarios"]:
        batch, newest, oldest, n_iter, seed, n_draws, cursor, full = (int(x) for x in g[f"{name}_meta"])
        rb = ReplayBuffer(tree, BatchSpec(T, B), size_T=size_T, replay_batch_size=batch,
                          newest_n_samples_invalid=newest, oldest_n_samples_invalid=oldest)
        rb.seed(seed)
        for _ in range(n_iter):
            rb.next_iteration()
        assert (rb._cursor, int(rb._full)) == (cursor, full)
        assert rb.capacity == size_T * B and rb.replay_batch_size == batch
        for d in range(n_draws):
            got = rb.sample_batch()
            assert set(got.keys()) == set(LEAVES) | {"t_index", "b_index"}
            for k in got:
                ref = g[f"{name}_d{d}_{k}"]
                out = got[k].cpu().numpy()
                assert out.dtype == ref.dtype and out.shape == ref.shape, (name, d, k)
                assert np.array_equal(out, ref), (name, d, k)


def test_host_mirror_tracks_sampler_writes():
    """Host-resident ring (shares memory with the sampler): next_iteration() uploads the T rows
    written since the last call, so samples always see the newest data."""
    from parllel_b200 import ArrayDict, BatchSpec
    from parllel_b200.replays import ReplayBuffer
    T, B, size_T = 4, 3, 16
    host = torch.zeros((size_T, B, 2))
    rb = ReplayBuffer(ArrayDict({"x": host}), BatchSpec(T, B), size_T=size_T, replay_batch_size=64)
    rb.seed(0)
    for it in range(6):  # wraps the ring
        lo = (it * T) % size_T
        host[lo:lo + T] = float(it + 1)
        rb.next_iteration()
        t_idx, b_idx = rb.sample_indices()
        got = rb.gather(t_idx, b_idx)["x"].cpu()
        assert torch.equal(got, host[torch.from_numpy(t_idx), torch.from_numpy(b_idx)])


@pytest.mark.parametrize("n", [1, 256, 16384])
def test_c3_ring_gather_bit_exact_vs_oracle(n):
    """BASELINE config C3: ~1M-transition ring (size_T 62464 x B 16), obs dim 64, act dim 8."""
    from parllel_b200 import ArrayDict, BatchSpec
    from parllel_b200.replays import ReplayBuffer
    size_T, B, T = 62464, 16, 128
    rng = np.random.default_rng(0)
    # observation is a view at row offset `padding` of a [size_T + 2, B, obs] base (sac.py:234-244)
    obs_base = torch.from_numpy(rng.standard_normal((size_T + 2, B, 64)).astype(np.float32)).cuda()
    host = {
        "observation": obs_base[1:-1],
        "next_observation": torch.from_numpy(rng.standard_normal((size_T, B, 64)).astype(np.float32)).cuda(),
        "action": torch.from_numpy(rng.standard_normal((size_T, B, 8)).astype(np.float32)).cuda(),
        "reward": torch.from_numpy(rng.standard_normal((size_T, B)).astype(np.float32)).cuda(),
        "terminated": torch.from_numpy(rng.random((size_T, B)) < 0.01).cuda(),
    }
    rb = ReplayBuffer(ArrayDict(host), BatchSpec(T, B), size_T=size_T, replay_batch_size=n,
                      oldest_n_samples_invalid=1)
    rb._full, rb._cursor = True, 3 * T
    rb.seed(0)
    ref_rng = O.Pcg64.from_numpy(np.random.default_rng(0))
    for _ in range(3):
        t_ref, b_ref = O.replay_sample_indices(ref_rng, size_T, B, 3 * T, True, 0, 1, n)
        t_idx, b_idx = rb.sample_indices()
        assert np.array_equal(t_idx, t_ref) and np.array_equal(b_idx, b_ref)
        got = rb.gather(t_idx, b_idx)
        for k, leaf in host.items():
            ref = O.gather_leaf(leaf.cpu().numpy(), t_ref, b_ref)
            assert np.array_equal(got[k].cpu().numpy(), ref), k


def test_gather_odd_row_sizes_and_alignments():
    """Rows of 1, 2, 3, 6, 12, 20 and 100 bytes; misaligned bases -> narrower chunks, same bytes."""
    from parllel_b200 import ArrayDict, BatchSpec
    from parllel_b200.replays import ReplayBuffer
    size_T, B = 32, 5
    rng = np.random.default_rng(1)
    leaves = {
        "u8": torch.from_numpy(rng.integers(0, 255, (size_T, B), dtype=np.uint8)),
        "i16": torch.from_numpy(rng.integers(-9, 9, (size_T, B), dtype=np.int16)),
        "u8x3": torch.from_numpy(rng.integers(0, 255, (size_T, B, 3), dtype=np.uint8)),
        "i16x3": torch.from_numpy(rng.integers(-9, 9, (size_T, B, 3), dtype=np.int16)),
        "f32x3": torch.from_numpy(rng.standard_normal((size_T, B, 3)).astype(np.float32)),
        "f32x5": torch.from_numpy(rng.standard_normal((size_T, B, 5)).astype(np.float32)),
        "f32x25": torch.from_numpy(rng.standard_normal((size_T, B, 5, 5)).astype(np.float32)),
        "f64": torch.from_numpy(rng.standard_normal((size_T, B))),
    }
    big = torch.from_numpy(rng.standard_normal((size_T, B, 9)).astype(np.float32)).cuda()
    dev = {k: v.cuda() for k, v in leaves.items()}
    dev["sliced"] = big[:, :, 1:5]  # 4-byte aligned only
    rb = ReplayBuffer(ArrayDict(dev), BatchSpec(4, B), size_T=size_T, replay_batch_size=77)
    rb._full = True
    rb.seed(5)
    t_idx, b_idx = rb.sample_indices()
    got = rb.gather(t_idx, b_idx)
    for k, v in dev.items():
        assert torch.equal(got[k].cpu(), v.cpu()[torch.from_numpy(t_idx), torch.from_numpy(b_idx)]), k


def test_k_fused_sampling_equals_sequential_calls(golden):
    """sample_batches(k) = k sample_batch() calls: same index stream, same bytes, one launch."""
    from parllel_b200 import BatchSpec
    from parllel_b200.replays import ReplayBuffer
    g = golden("replay")
    tree, T, B, size_T = _ring_tree(g, "cuda")
    a = ReplayBuffer(tree, BatchSpec(T, B), size_T=size_T, replay_batch_size=33, oldest_n_samples_invalid=1)
    b = ReplayBuffer(tree, BatchSpec(T, B), size_T=size_T, replay_batch_size=33, oldest_n_samples_invalid=1)
    for rb in (a, b):
        rb._full, rb._cursor = True, 24
        rb.seed(11)
    fused = a.sample_batches(5)
    for i in range(5):
        single = b.sample_batch()
        for k in single:
            assert torch.equal(fused[k][i], single[k]), (i, k)
    # the generators stay aligned afterwards
    assert np.array_equal(a.sample_indices()[0], b.sample_indices()[0])


@pytest.mark.parametrize("residency", ["cuda", "host"])
def test_batched_dataloader_matches_reference_golden(golden, residency):
    from parllel_b200 import ArrayDict, BatchSpec
    from parllel_b200.replays import BatchedDataLoader
    g = golden("dataloader")
    T, B = (int(x) for x in g["meta"])
    leaves = {k: torch.from_numpy(g[f"in_{k}"]) for k in ("observation", "advantage", "return_", "action")}
    tt, bb = np.meshgrid(np.arange(T), np.arange(B), indexing="ij")
    leaves["t_index"] = torch.from_numpy(tt.copy())
    leaves["b_index"] = torch.from_numpy(bb.copy())
    if residency == "cuda":
        leaves = {k: v.cuda() for k, v in leaves.items()}
    for name, n_batches, drop_last, shuffle in [("even", 4, True, True), ("ragged_drop", 5, True, True),
                                                ("ragged_keep", 5, False, True), ("noshuffle", 3, True, False)]:
        dl = BatchedDataLoader(ArrayDict(leaves), BatchSpec(T, B), n_batches=n_batches, drop_last=drop_last,
                               shuffle=shuffle)
        dl.seed(9)
        n = 0
        for epoch in range(2):
            for batch in dl.batches():
                for k in batch:
                    ref = g[f"{name}_{n}_{k}"]
                    got = batch[k].cpu().numpy()
                    assert got.shape == ref.shape and got.dtype == ref.dtype, (name, n, k)
                    assert np.array_equal(got, ref), (name, n, k)
                n += 1
        assert n == int(g[f"{name}_meta"][3])
        assert len(dl) == T * B


def test_batched_dataloader_recurrent_and_batch_only_fields():
    """Recurrent mode gathers whole [T] columns; batch_only_fields are indexed by env only
    (batched_dataloader.py:75-87,104-106) -- checked against torch advanced indexing."""
    from parllel_b200 import ArrayDict, BatchSpec
    from parllel_b200.replays import BatchedDataLoader
    T, B = 12, 10
    rng = np.random.default_rng(3)
    tree = ArrayDict({
        "observation": torch.from_numpy(rng.standard_normal((T, B, 7)).astype(np.float32)).cuda(),
        "valid": torch.from_numpy(rng.random((T, B)) < 0.8).cuda(),
        "initial_rnn_state": torch.from_numpy(rng.standard_normal((B, 3)).astype(np.float32)).cuda(),
    })
    dl = BatchedDataLoader(tree, BatchSpec(T, B), n_batches=3, recurrent=True, batch_only_fields=["initial_rnn_state"])
    dl.seed(1)
    ref_rng = np.random.default_rng(1)
    order = np.arange(B, dtype=np.int32)
    ref_rng.shuffle(order)
    batches = list(dl.batches())
    assert len(batches) == 3
    for i, batch in enumerate(batches):
        idx = torch.from_numpy(order[i * 3:(i + 1) * 3].astype(np.int64)).cuda()
        assert torch.equal(batch["observation"], tree["observation"][:, idx])
        assert torch.equal(batch["valid"], tree["valid"][:, idx])
        assert torch.equal(batch["initial_rnn_state"], tree["initial_rnn_state"][idx])


@pytest.mark.parametrize("size_T,B,cursor,full,newest,oldest,n,k", [
    (62464, 16, 384, True, 0, 1, 16384, 1),     # C3
    (62464, 16, 384, True, 0, 1, 256, 7),       # K-fused minibatches
    (64, 4, 24, True, 2, 3, 33, 5),             # odd sizes: the half-word buffer crosses calls
    (64, 4, 40, False, 1, 0, 7, 3),             # ring not full yet
    (2**31 + 11, 3, 5, True, 0, 0, 5000, 2),    # ~half of all words rejected
    (1000003, 1, 17, True, 0, 0, 1025, 3),      # B == 1 draws nothing (numpy returns zeros)
    (4, 2**32, 3, False, 0, 0, 300, 2),         # full 32-bit range: raw words
])
def test_device_index_generation_is_bit_exact_with_numpy(size_T, B, cursor, full, newest, oldest, n, k):
    """plb_replay_sample_indices continues numpy's PCG64 stream bit for bit, state included."""
    from parllel_b200 import _lib
    gen = np.random.default_rng(123)
    gen.integers(0, 10, size=3)  # leave a buffered half-word behind, like a generator in use
    st = gen.bit_generator.state
    s, inc, m = st["state"]["state"], st["state"]["inc"], (1 << 64) - 1
    words = np.array([s >> 64, s & m, inc >> 64, inc & m, (st["has_uint32"] & 1) | (int(st["uinteger"]) << 32)],
                     dtype=np.uint64)
    state = torch.from_numpy(words.view(np.int64).copy()).cuda()
    t = torch.empty((k, n), dtype=torch.int64, device="cuda")
    b = torch.empty((k, n), dtype=torch.int64, device="cuda")
    lib = _lib.load()
    for rnd in range(2):  # the second round starts from the state the kernel left behind
        _lib.check(lib.plb_replay_sample_indices(state.data_ptr(), size_T, B, cursor, int(full), newest, oldest, n, k,
                                                 t.data_ptr(), b.data_ptr(), torch.cuda.current_stream().cuda_stream))
        for kk in range(k):
            if full:
                t_ref = (gen.integers(0, size_T - oldest - newest, size=(n,)) + cursor + oldest) % size_T
            else:
                t_ref = gen.integers(0, cursor - newest, size=(n,))
            b_ref = gen.integers(0, B, size=(n,))
            assert np.array_equal(t[kk].cpu().numpy(), t_ref), (rnd, kk)
            assert np.array_equal(b[kk].cpu().numpy(), b_ref), (rnd, kk)
    st = gen.bit_generator.state
    got = state.cpu().numpy().view(np.uint64)
    assert int(got[0]) == st["state"]["state"] >> 64 and int(got[1]) == st["state"]["state"] & m
    assert int(got[4]) & 1 == st["has_uint32"]
    if st["has_uint32"]:
        assert int(got[4]) >> 32 == st["uinteger"]


def test_replay_buffer_device_rng_matches_host_rng(golden):
    from parllel_b200 import BatchSpec
    from parllel_b200.replays import ReplayBuffer
    g = golden("replay")
    tree, T, B, size_T = _ring_tree(g, "cuda")
    host = ReplayBuffer(tree, BatchSpec(T, B), size_T=size_T, replay_batch_size=33, oldest_n_samples_invalid=1)
    dev = ReplayBuffer(tree, BatchSpec(T, B), size_T=size_T, replay_batch_size=33, oldest_n_samples_invalid=1,
                       device_rng=True)
    for rb in (host, dev):
        rb._full, rb._cursor = True, 24
        rb.seed(5)
    for _ in range(4):
        a, b = host.sample_batch(), dev.sample_batch()
        for key in a:
            assert torch.equal(a[key], b[key]), key
    fa, fb = host.sample_batches(3), dev.sample_batches(3)
    for key in fa:
        assert torch.equal(fa[key], fb[key]), key
