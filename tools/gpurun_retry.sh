#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged).  usage: gpurun_retry.sh LOG [gpurun args...]
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
