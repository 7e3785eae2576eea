#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>'   — retries while the pod has no free GPU slot (exit code 3)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
