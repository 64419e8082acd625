#!/bin/bash
# round-2 (second session) final single-GPU evidence: tests, smoke, the contract bench, the reference arm, ncu captures
# (each ncu capture only after the same command has run clean without ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r3_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r3_bench_reference.json 2> gpurun_out/r3_bench_reference.err; echo "reference arm rc=$?"
timeout 60 python tools/pcie_probe.py 2>&1 | tail -3 > gpurun_out/r3_pcie.log; cat gpurun_out/r3_pcie.log
NCU="ncu --set full --clock-control none --import-source on -f"
{
python tools/run_kernels.py step --T 1024 --B 4096 --reps 3 --resident 0 > /dev/null 2>&1 || echo "FAILED plain run (step)"
$NCU -k regex:scan_pipelined -s 2 -c 1 -o gpurun_out/r3_scan_stats python tools/run_kernels.py step --T 1024 --B 4096 --reps 3 --resident 0 | tail -2
$NCU -k regex:normalize_advantage_partials -s 2 -c 1 -o gpurun_out/r3_normadv_partials python tools/run_kernels.py step --T 1024 --B 4096 --reps 3 --resident 0 | tail -2
python bench.py --steps 20 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/r3_bench_plain.json 2> gpurun_out/r3_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 600 --csv --log-file gpurun_out/r3_launches_bench.csv python bench.py --steps 20 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/r3_ncu_bench.log 2>&1
ls -la gpurun_out/r3_*.ncu-rep gpurun_out/r3_launches_bench.csv
} > gpurun_out/r3_ncu.log 2>&1
tail -8 gpurun_out/r3_ncu.log
