#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers "busy" (exit code 3)
t=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  echo "[gpu_retry] busy (attempt $attempt), sleeping 90 s"
  sleep 90
done
exit 3
