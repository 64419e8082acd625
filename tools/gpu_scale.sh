#!/bin/bash
# scaling run of the contract command on one 8-GPU box: N = 2, 4, 8 (default route); N = 1 comes from the 1-GPU runs
mkdir -p gpurun_out
for W in ${1:-2 4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $W --steps 1000 --warmup 10 > gpurun_out/scale_n${W}.json 2> gpurun_out/scale_n${W}.err; echo "bench N=$W rc=$?"
done
python - <<PY
import json
for W in "${1:-2 4 8}".split():
    try:
        for line in open("gpurun_out/scale_n%s.json" % W):
            if line.startswith("{"):
                d = json.loads(line); c4 = d["c4_sharded"]
                print("N=%s: C1 step %.2f us = %.1f G tr/s, windows %s; C4 %.2f us (local %.2f, wait %.2f); e2e %.3g tr/s; parity %.3f/%.3f" % (
                    W, d["ms_per_step"] * 1e3, d["value"] / 1e9, [round(x * 1e3, 2) for x in d["windows_ms_per_step"]], c4["us_per_step"],
                    c4.get("local_only_us_per_step", 0), c4.get("exchange_wait_us_per_step", 0), d["e2e"]["value"], d["parity"]["c1_max_violation"], d["parity"]["c4_max_violation"]))
    except Exception as e:
        print(W, "failed", e)
PY
