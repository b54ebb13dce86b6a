#!/bin/bash
# usage: gpurun_retry.sh <gpus> <timeout> <script>   - retries while the pool answers "transient"
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --gpus $1 --timeout $2 -- "bash $3" 2>&1)
  echo "$out" | tail -15
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  break
done
