#!/bin/bash
# scripts/gpuretry_n.sh NGPUS TIMEOUT 'command'
n=$1; t=$2; shift 2
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --gpus $n --timeout $t -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
