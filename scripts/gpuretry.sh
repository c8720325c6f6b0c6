#!/bin/bash
# scripts/gpuretry.sh TIMEOUT 'command' : retries gpurun while the pod answers "busy" (exit 3 / transient)
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
