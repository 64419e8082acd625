#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> <command...>   -- retries while the pod answers busy (rc 3 / transient)
T=$1; G=$2; shift 2
for attempt in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  echo "$OUT" | tail -120
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then echo "[retry] attempt $attempt busy; sleeping 150 s"; sleep 150; continue; fi
  break
done
