#!/bin/bash
# usage: profiles/gpu_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (nothing charged)
t=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  break
done
