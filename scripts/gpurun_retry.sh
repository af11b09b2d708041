#!/bin/bash
# Development: gpurun with retries while the pod has no free GPU slot (nothing is charged for those attempts).
# usage: scripts/gpurun_retry.sh [--gpus N] <timeout-seconds> '<command>'
G=""
if [ "$1" == "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\b"; then sleep 100; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
