"""CPU (gloo, world_size 2) tests of the multi-rank host logic: column sharding and the all-gather of
the per-rank (count, mean, M2) moments.  The rank-ordered merge itself is a CUDA kernel and is
covered by tests/test_gpu_distributed.py; here its maths is checked with a numpy restatement."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _chan_merge(parts):
    n, mean, m2 = 0.0, 0.0, 0.0
    for pn, pm, pm2 in parts:
        if pn == 0:
            continue
        if n == 0:
            n, mean, m2 = pn, pm, pm2
            continue
        tot = n + pn
        d = pm - mean
        mean, m2, n = mean + d * pn / tot, m2 + pm2 + d * d * n * pn / tot, tot
    return n, mean, m2


def _worker(rank, world, port, B, seed, out_queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from parllel_b200.distributed import all_gather_moments, column_shard
    rng = np.random.default_rng(seed)
    full = rng.standard_normal((32, B)) * 3 + 1  # every rank generates the same global batch
    sl = column_shard(B, world, rank)
    mine = full[:, sl]
    local = torch.tensor([mine.size, mine.mean(), ((mine - mine.mean()) ** 2).sum()], dtype=torch.float64)
    gathered = all_gather_moments(local)
    assert gathered.shape == (world, 3)
    assert torch.equal(gathered[rank], local)
    n, mean, m2 = _chan_merge([tuple(row.tolist()) for row in gathered])
    out_queue.put((rank, sl.start, sl.stop, n, mean, m2, full.size, float(full.mean()), float(full.var())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [16, 17])
def test_sharded_moments_all_gather_and_merge_gloo(B):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile [0, B) without overlap
    assert results[0][1] == 0 and results[0][2] == results[1][1] and results[1][2] == B
    for rank, lo, hi, n, mean, m2, size, ref_mean, ref_var in results:
        assert n == size
        np.testing.assert_allclose(mean, ref_mean, rtol=1e-13)
        np.testing.assert_allclose(m2 / n, ref_var, rtol=1e-12)
    # every rank derived the same statistics bit for bit (same inputs, same merge order)
    assert results[0][3:6] == results[1][3:6]


def test_column_shard_partitions():
    from parllel_b200.distributed import column_shard
    for B in (1, 7, 8, 65536, 4097):
        for world in (1, 2, 3, 8):
            edges = [column_shard(B, world, r) for r in range(world)]
            assert edges[0].start == 0 and edges[-1].stop == B
            assert all(a.stop == b.start for a, b in zip(edges, edges[1:]))
            sizes = [e.stop - e.start for e in edges]
            assert max(sizes) - min(sizes) <= 1
