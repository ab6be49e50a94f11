#!/bin/bash
# usage: tools/gpu_retry.sh [--gpus N] TIMEOUT 'command'   -- retries while the pod answers "busy" (exit 3 / transient)
G=""
if [ "$1" == "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout $T -- "$@" 2>&1); rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
