#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout> <command...>   — retries a gpurun call while the pod answers "busy" (exit 3)
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' /root/repo/gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 60
done
exit 3
