#!/bin/bash
# like gpurun_retry.sh, for a 2-GPU call
T=$1; shift
for i in $(seq 1 20); do
  out=$(gpurun --gpus 2 --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out" | tail -80
  exit 0
done
echo "gave up: pod busy"
