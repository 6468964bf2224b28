#!/bin/bash
# usage: gpurun_retry.sh <timeout> <logfile> <command...>
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG; then sleep 120; continue; fi
  break
done
echo "DONE rc=$rc" >> $LOG
