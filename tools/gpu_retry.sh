#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>' : retries while the pod answers busy (exit code 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
