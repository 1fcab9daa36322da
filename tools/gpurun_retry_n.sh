#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <log> <timeout> <command...>   (retries while the pod has no slot)
G=$1; shift; LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
