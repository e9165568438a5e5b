#!/bin/bash
# local helper: run a gpurun call, retrying while the pool answers "busy" (exit 3: nothing charged)
#   tools/gpurun_retry.sh <log> <timeout> <command...>
log=$1; shift; to=$1; shift
for try in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to ${GPURUN_FLAGS} -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
