#!/bin/bash
mkdir -p gpurun_out
{
timeout 120 python tools/resident_timeline.py
timeout 120 python tools/resident_timeline.py 15=1
timeout 120 python tools/resident_timeline.py 14=1
timeout 120 python tools/resident_timeline.py 5=4
timeout 120 python tools/resident_timeline.py --T 256 --B 8192
for args in "--resident 1" "--resident 0"; do
  timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 24 --reps 15 $args
done
PLB_TUNING=15=1 timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 24 --reps 15
PLB_TUNING=14=1 timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 24 --reps 15
timeout 120 python tools/run_kernels.py step --T 256 --B 8192 --inner 24 --reps 15
timeout 300 python -m pytest tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/r2b.log 2>&1
cat gpurun_out/r2b.log
