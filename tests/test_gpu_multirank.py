"""Multi-GPU parity on hardware: spawns one process per GPU (torchrun) running tests/multirank_worker.py,
which checks every rank's shard against the CPU oracle evaluated on the GLOBAL batch, bit-identical
statistics across ranks, NCCL route == in-kernel routes (record per rank, resident kernel), and protocol agreement
on uneven shards.
Skipped where fewer GPUs are visible; `gpurun --gpus N -- python -m pytest tests/test_gpu_multirank.py -m gpu`
runs it, and the JSON reports it writes are kept under profiles/."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_transforms_match_the_oracle_on_the_global_batch(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    log = os.path.join(ROOT, "gpurun_out", f"multirank_w{world}.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multirank_worker.py")]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PLB_MULTIRANK_LOG=log)
    proc = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (proc.stdout + proc.stderr)[-6000:]
    assert proc.returncode == 0, tail
    assert f"MULTIRANK OK world={world}" in proc.stdout, tail
