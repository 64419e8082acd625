#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_resident.py tests/test_gpu_replay.py tests/test_gpu_transforms.py tests/test_gpu_scan.py -x -q -m gpu 2>&1 | tail -5
timeout 120 python tools/resident_timeline.py
timeout 120 python tools/resident_timeline.py 15=1
timeout 120 python tools/resident_timeline.py 5=4
timeout 120 python tools/resident_timeline.py --T 256 --B 8192
for args in "--resident 1" "--resident 0"; do
  timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 24 --reps 15 $args
  timeout 120 python tools/run_kernels.py step --T 256 --B 8192 --inner 24 --reps 15 $args
  timeout 120 python tools/run_kernels.py chain --T 1024 --B 4096 --inner 24 --reps 15 $args
done
timeout 300 python tools/time_replay.py default 16=1
} > gpurun_out/r2d.log 2>&1
cat gpurun_out/r2d.log
