#!/bin/bash
# usage: gpu_retry.sh <gpurun args...>   -- retries while the pod answers "transient"
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 1
