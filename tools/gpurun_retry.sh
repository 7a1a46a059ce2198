#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient): nothing is charged for those.
# usage: tools/gpurun_retry.sh <logfile> [gpurun args...] -- '<command>'
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
