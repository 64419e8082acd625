#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 120 python tools/run_kernels.py gae --T 1024 --B 4096 --inner 24 --reps 15
timeout 120 python tools/run_kernels.py gae --T 256 --B 65536 --inner 24 --reps 15
timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 24 --reps 15
timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 1 --reps 50
timeout 120 python tools/run_kernels.py chain --T 1024 --B 4096 --inner 24 --reps 15
timeout 600 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2g_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','parity_checked','gpu_launches','kernels_per_step','windows_ms_per_step')}, d['parity'])
print(d['roofline']['kernel'][:60], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline']['step_frac'])
print(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['single_batch_latency_ms'])
"
} > gpurun_out/r2g.log 2>&1
cat gpurun_out/r2g.log
