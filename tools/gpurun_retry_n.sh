#!/bin/bash
# usage: gpurun_retry_n.sh <gpus> <timeout_s> <command...>   — like gpurun_retry.sh for a multi-GPU box
N=$1; T=$2; shift; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 90
done
cat /tmp/gpurun_last.log | tail -60
exit $rc
