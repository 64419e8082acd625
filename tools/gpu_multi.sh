#!/bin/bash
# usage: bash tools/gpu_multi.sh N ["worlds"]   -- multi-rank parity test + bench on N GPUs of one box
N=$1
WORLDS=${2:-"2 4 8"}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi${N}_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu -x > gpurun_out/multi${N}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi${N}_pytest.log
tail -25 gpurun_out/multi${N}_pytest.log
for W in $WORLDS; do
  if [ $W -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $W --steps 1000 --warmup 10 > gpurun_out/bench_n${W}.json 2> gpurun_out/bench_n${W}.err; echo "bench N=$W rc=$?"
    tail -c 1500 gpurun_out/bench_n${W}.json; tail -5 gpurun_out/bench_n${W}.err
    for R in 1; do
      PLB_TUNING=17=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $W --steps 500 --warmup 10 --windows 3 > gpurun_out/bench_n${W}_route$R.json 2> gpurun_out/bench_n${W}_route$R.err; echo "bench N=$W route $R rc=$?"
    done
    python - <<PY
import json
for tag in ("", "_route1"):
    try:
        for line in open("gpurun_out/bench_n${W}%s.json" % tag):
            if line.startswith("{"):
                d = json.loads(line)
                c4 = d["c4_sharded"]
                print("N=${W} route%s: C1 step %.2f us (kernels/step %d, dominant kernel %.2f us), parity %s; C4 %.2f us (local %.2f, wait %.2f), e2e %.3g tr/s" % (
                    tag or "(default)", d["ms_per_step"] * 1e3, d["kernels_per_step"], d["roofline"]["avg_launch_us"], d["parity"],
                    c4["us_per_step"], c4.get("local_only_us_per_step", 0), c4.get("exchange_wait_us_per_step", 0), d["e2e"]["value"]))
    except Exception as e:
        print(tag, "failed", e)
PY
  fi
done
