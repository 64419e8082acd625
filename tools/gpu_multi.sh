#!/bin/bash
# usage: bash tools/gpu_multi.sh N   -- multi-rank parity test + bench on N GPUs of one box
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi${N}_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu -x > gpurun_out/multi${N}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi${N}_pytest.log
tail -25 gpurun_out/multi${N}_pytest.log
for W in 2 4 8; do
  if [ $W -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $W --steps 1000 --warmup 10 > gpurun_out/bench_n${W}.json 2> gpurun_out/bench_n${W}.err; echo "bench N=$W rc=$?"
    tail -c 3000 gpurun_out/bench_n${W}.json; tail -5 gpurun_out/bench_n${W}.err
  fi
done
