#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).  Usage:
#   tools/gpurun_retry.sh <log file> [gpurun options] -- '<command>'
log=$1; shift
for attempt in $(seq 1 20); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 150
done
exit 3
