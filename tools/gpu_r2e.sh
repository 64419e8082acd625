#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/resident_timeline.py
timeout 120 python tools/resident_timeline.py --T 256 --B 8192
timeout 120 python tools/resident_timeline.py --T 256 --B 32768
for args in "--resident 1" "--resident 0"; do
  timeout 120 python tools/run_kernels.py step --T 1024 --B 4096 --inner 24 --reps 15 $args
  timeout 120 python tools/run_kernels.py step --T 256 --B 8192 --inner 24 --reps 15 $args
  timeout 120 python tools/run_kernels.py step --T 256 --B 32768 --inner 24 --reps 15 $args
done
timeout 600 python bench.py --steps 500 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2e_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','parity_checked','parity','gpu_launches','kernels_per_step','windows_ms_per_step')})
print(d['roofline']['kernel'][:60], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline']['step_frac'])
for r in d['roofline'].get('secondary',[]): print(r['name'], round(r['avg_launch_us'],2), round(r['frac'],3))
print(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['single_batch_latency_ms'])
print(json.dumps(d['e2e'].get('secondary'),indent=0)[:1500])
print(d.get('cpu_baseline'))
print(json.dumps(d.get('cpu_baseline_secondary'))[:1500])
print(d['c4_sharded'])
"
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 | cut -c1-400
} > gpurun_out/r2e.log 2>&1
cat gpurun_out/r2e.log
