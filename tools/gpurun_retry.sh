#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'   (retries while the pod answers busy/transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gpurun: still busy after 40 tries"; exit 3
