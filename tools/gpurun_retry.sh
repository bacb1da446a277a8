#!/bin/bash
# tools/gpurun_retry.sh <out-file> <gpurun args...>: retries while the pod answers "transient"/busy (nothing is charged then)
out=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if grep -q "status=transient\|rc=3\|status=busy" "$out"; then sleep 90; else break; fi
done
