#!/bin/bash
# usage: [GPURUN_FLAGS="--gpus 2"] tools/gpurun_retry.sh <timeout> <logfile> <command...>
# retries while the pod answers busy (exit 3 / transient)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
