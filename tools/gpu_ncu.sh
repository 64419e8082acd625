#!/bin/bash
# round 2 ncu evidence (each capture only after its command has run clean without ncu)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
run() { echo "== $*"; "$@" > /dev/null 2>&1 || echo "FAILED plain run: $*"; }
{
# two-kernel route of the step (what bench.py runs at N = 1) and the resident kernel (what it runs at N > 1)
run python tools/run_kernels.py step --T 1024 --B 4096 --reps 3 --resident 0
$NCU -k regex:scan_pipelined -s 2 -c 1 -o gpurun_out/r2_scan_stats python tools/run_kernels.py step --T 1024 --B 4096 --reps 3 --resident 0 | tail -2
$NCU -k regex:normalize_advantage_partials -s 2 -c 1 -o gpurun_out/r2_normadv_partials python tools/run_kernels.py step --T 1024 --B 4096 --reps 3 --resident 0 | tail -2
run python tools/run_kernels.py step --T 1024 --B 4096 --reps 3 --resident 2
$NCU -k regex:scan_pipelined -s 2 -c 1 -o gpurun_out/r2_resident python tools/run_kernels.py step --T 1024 --B 4096 --reps 3 --resident 2 | tail -2
run python tools/run_kernels.py past --T 1024 --B 4096 --reps 3
$NCU -k regex:scan_pipelined -s 2 -c 1 -o gpurun_out/r2_past_return python tools/run_kernels.py past --T 1024 --B 4096 --reps 3 | tail -2
run python tools/run_kernels.py gather --n 16384 --reps 3
$NCU -k regex:gather_tree -s 2 -c 1 -o gpurun_out/r2_gather_n16384 python tools/run_kernels.py gather --n 16384 --reps 3 | tail -2
run python tools/time_replay.py default
$NCU -k regex:replay_indices_cluster -s 40 -c 1 -o gpurun_out/r2_rng_cluster python tools/time_replay.py default | tail -2
run python tools/run_kernels.py norm_obs --reps 3
$NCU -k regex:obs_step -s 2 -c 1 -o gpurun_out/r2_obs python tools/run_kernels.py norm_obs --reps 3 | tail -2
# per-kernel durations of the three sharded routes run on one GPU (experiments; ncu serialises the kernels)
for R in 2 3; do
  PLB_TUNING=17=$R ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20 -c 12 --csv --log-file gpurun_out/r2_launches_route$R.csv python tools/run_kernels.py step --T 1024 --B 4096 --reps 20 > /dev/null 2>&1
  echo "route $R:"; grep -v "^==" gpurun_out/r2_launches_route$R.csv | awk -F'","' 'NR>1 {print $5, $NF}' | cut -c1-160 | tail -6
done
# launch list of the contract command
python bench.py --steps 20 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 20 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_bench.csv
} > gpurun_out/r2_ncu.log 2>&1
tail -40 gpurun_out/r2_ncu.log
