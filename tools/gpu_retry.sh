#!/bin/bash
# usage: tools/gpu_retry.sh <log> <gpurun args...>   -- retries while the pod answers "busy" (exit 3), nothing is charged for those
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then break; fi
  sleep 60
done
echo "gpu_retry exit $rc" >> "$log"
