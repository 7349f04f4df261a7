#!/bin/bash
# gpurun with retries on "no box or slot free right now" (exit 3): tools/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
