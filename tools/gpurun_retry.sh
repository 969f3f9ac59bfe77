#!/bin/bash
# tools/gpurun_retry.sh LOGFILE TIMEOUT CMD... : run a gpurun call, retrying while the pod answers busy (exit 3 / transient)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then break; fi
  sleep 60
done
echo finished >> $LOG
