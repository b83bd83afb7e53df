#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args...]
# Retries while the pod answers "no box or slot free" (rc 3) or while an earlier call of this repo is still in flight.
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out"
  if [ $rc -eq 3 ]; then sleep 60; continue; fi
  if [ $rc -eq 2 ] && echo "$out" | grep -q "already running"; then sleep 45; continue; fi
  exit $rc
done
exit 3
