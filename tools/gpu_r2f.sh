#!/bin/bash
mkdir -p gpurun_out
{
for tun in "17=0" "17=2" "12=2"; do
  echo "== PLB_TUNING=$tun"
  PLB_TUNING=$tun timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 24 --reps 15
  PLB_TUNING=$tun timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 1 --reps 30
  PLB_TUNING=$tun timeout 120 python tools/run_kernels.py step --T 256 --B 32768 --inner 24 --reps 15
done
timeout 300 python -m pytest tests/test_gpu_resident.py tests/test_gpu_replay.py -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/r2f.log 2>&1
cat gpurun_out/r2f.log
