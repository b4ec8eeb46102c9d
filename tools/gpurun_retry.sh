#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers "busy" (nothing is charged for those)
T=$1; shift
for i in $(seq 1 20); do
  out=$(gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out" | tail -80
  exit 0
done
echo "gave up: pod busy"
