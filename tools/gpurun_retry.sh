#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> <command...>   — retries while the pod answers "busy" (nothing charged)
T=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 90
done
cat /tmp/gpurun_last.log | tail -40
exit $rc
