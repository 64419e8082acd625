"""GPU tests of the multi-rank statistics exchange.  With one GPU the ranks are emulated as column
shards processed one after the other (never as concurrently spinning kernels); the NCCL path itself
is exercised by `bench.py --gpus N` under torchrun."""
import numpy as np
import pytest
import torch

from tests.util import assert_within_tolerance, make_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_rank_ordered_merge_equals_whole_batch(world):
    from parllel_b200 import _lib
    from parllel_b200.distributed import column_shard, merge_gathered_moments
    rng = np.random.default_rng(world)
    T, B = 64, 256
    x = torch.from_numpy((rng.standard_normal((T, B)) * 2 + 7).astype(np.float32)).cuda()
    lib = _lib.load()
    ws = torch.zeros(lib.plb_batch_moments_workspace_bytes(), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    parts = torch.zeros((world, 3), dtype=torch.float64, device="cuda")
    for r in range(world):
        sl = column_shard(B, world, r)
        shard = x[:, sl]
        _lib.check(lib.plb_batch_moments_f32(shard.data_ptr(), B, T, sl.stop - sl.start, 0, 0,
                                             parts[r].data_ptr(), ws.data_ptr(), ws.numel(), s))
    merged = torch.zeros(3, dtype=torch.float64, device="cuda")
    merge_gathered_moments(parts, 1, 0, merged)
    whole = torch.zeros(3, dtype=torch.float64, device="cuda")
    _lib.check(lib.plb_batch_moments_f32(x.data_ptr(), B, T, B, 0, 0, whole.data_ptr(), ws.data_ptr(), ws.numel(), s))
    n, mean, m2 = merged.cpu().numpy()
    ref = x.cpu().numpy().astype(np.float64)
    assert n == ref.size
    np.testing.assert_allclose(mean, ref.mean(), rtol=1e-13)
    np.testing.assert_allclose(m2 / n, ref.var(), rtol=1e-12)
    np.testing.assert_allclose(merged.cpu().numpy(), whole.cpu().numpy(), rtol=1e-12)


def test_column_layout_merge():
    from parllel_b200.distributed import merge_gathered_moments
    rng = np.random.default_rng(0)
    world, F, rows = 3, 37, 50
    data = [rng.standard_normal((rows + r, F)) for r in range(world)]
    parts = torch.zeros((world, 1 + 2 * F), dtype=torch.float64)
    for r, d in enumerate(data):
        parts[r, 0] = d.shape[0]
        parts[r, 1:1 + F] = torch.from_numpy(d.mean(0))
        parts[r, 1 + F:] = torch.from_numpy(((d - d.mean(0)) ** 2).sum(0))
    parts = parts.cuda()
    out = torch.zeros(1 + 2 * F, dtype=torch.float64, device="cuda")
    merge_gathered_moments(parts, F, 1, out)
    allv = np.concatenate(data)
    o = out.cpu().numpy()
    assert o[0] == allv.shape[0]
    np.testing.assert_allclose(o[1:1 + F], allv.mean(0), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(o[1 + F:] / o[0], allv.var(0), rtol=1e-11)


def test_sharded_pipeline_emulated_on_one_gpu():
    """C4-style sharding emulated: each 'rank' runs phases A/B on its column block, moments are merged
    in rank order, then phase C; the result must equal the unsharded chain on the whole batch."""
    import ctypes
    from parllel_b200 import _lib
    from parllel_b200.distributed import column_shard, merge_gathered_moments
    from parllel_b200.transforms import EstimateAdvantage, FusedBatchTransforms, NormalizeAdvantage
    T, B, world = 256, 512, 4
    batch = make_batch(np.random.default_rng(2), T, B, 0.01)
    dev = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}

    def tree(sl):
        return {"reward": dev["reward"][:, sl].contiguous(), "done": dev["done"][:, sl].contiguous(),
                "agent_info": {"value": dev["value"][:, sl].contiguous()},
                "bootstrap_value": dev["bootstrap_value"][sl].contiguous(),
                "advantage": torch.zeros(T, sl.stop - sl.start, device="cuda"),
                "return_": torch.zeros(T, sl.stop - sl.start, device="cuda")}

    whole = tree(slice(0, B))
    FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(whole)], use_cuda_graph=False)(whole)
    shards = [tree(column_shard(B, world, r)) for r in range(world)]
    pipes = [FusedBatchTransforms([EstimateAdvantage(0.99, 0.95), NormalizeAdvantage(t)], use_cuda_graph=False)
             for t in shards]
    for pipe, t in zip(pipes, shards):
        pipe(t, phases=_lib.PHASE_A | _lib.PHASE_B)
    gathered = torch.stack([pipe._moments[3:6] for pipe in pipes])
    merged = torch.zeros(3, dtype=torch.float64, device="cuda")
    merge_gathered_moments(gathered, 1, 0, merged)
    for pipe, t in zip(pipes, shards):
        pipe._moments[3:6] = merged
        pipe(t, phases=_lib.PHASE_C)
    got = torch.cat([t["advantage"] for t in shards], dim=1)
    assert_within_tolerance(got.cpu().numpy(), whole["advantage"].cpu().numpy(), rel=2e-6, what="sharded advantage")
