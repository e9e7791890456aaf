#!/bin/bash
# usage: gpurun_retry.sh <logfile> <timeout> <command...>   - retries while the pod answers busy (exit 3 / transient)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpurun_retry done rc=$rc attempt=$attempt" >> $log
