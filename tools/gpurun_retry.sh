#!/bin/bash
# gpurun with retries while the pod has no slot (exit code 3 = transient, nothing charged).
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
