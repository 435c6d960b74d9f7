#!/bin/bash
# usage: scripts/gpu_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
log=$1; shift
for attempt in $(seq 1 30); do
  gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 90
done
exit 3
