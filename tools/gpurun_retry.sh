#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'
# Retries ONLY while the pod answers busy (exit code 3 / status=transient: nothing ran, nothing was charged).  Anything else --
# in particular a run that lost its box -- is final: resubmitting a command that killed a box is how round 2 lost its GPU access
# (three strikes from one retried command).
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if [ $rc -eq 3 ] || echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gpurun: still busy after 40 tries"; exit 3
