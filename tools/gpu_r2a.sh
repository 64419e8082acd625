#!/bin/bash
# round 2, first GPU session: parity of the resident kernel + timing sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_resident.py -x -q -m gpu > gpurun_out/r2a_resident.log 2>&1; echo "resident rc=$?" >> gpurun_out/r2a_resident.log
tail -15 gpurun_out/r2a_resident.log
timeout 900 python -m pytest tests/test_gpu_transforms.py tests/test_gpu_scan.py tests/test_gpu_distributed.py tests/test_gpu_replay.py -q -m gpu > gpurun_out/r2a_rest.log 2>&1; echo "rest rc=$?" >> gpurun_out/r2a_rest.log
tail -15 gpurun_out/r2a_rest.log
{
for args in "--resident 1" "--resident 1 --debug 4" "--resident 0" "--resident 1 --stages 3" "--resident 1 --stages 2"; do
  timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 24 --reps 15 $args
done
for B in 8192 16384 32768 65536; do
  timeout 120 python tools/run_kernels.py step --T 256 --B $B --inner 24 --reps 15 --resident 1
  timeout 120 python tools/run_kernels.py step --T 256 --B $B --inner 24 --reps 15 --resident 0
done
timeout 120 python tools/run_kernels.py chain --T 1024 --B 4096 --inner 24 --reps 15 --resident 1
timeout 120 python tools/run_kernels.py chain --T 1024 --B 4096 --inner 24 --reps 15 --resident 0
timeout 120 python tools/run_kernels.py gae --T 1024 --B 4096 --inner 24 --reps 15
timeout 120 python tools/run_kernels.py past --T 1024 --B 4096 --inner 24 --reps 15
} > gpurun_out/r2a_timing.log 2>&1
cat gpurun_out/r2a_timing.log
timeout 300 python tools/time_replay.py default 13=0 > gpurun_out/r2a_replay.log 2>&1
cat gpurun_out/r2a_replay.log
