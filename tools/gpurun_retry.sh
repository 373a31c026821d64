#!/bin/bash
# usage: tools/gpurun_retry.sh <log> [gpurun args...] -- command ; retries while the pod answers "busy" (exit code 3)
LOG=$1; shift
for i in 1 2 3 4 5 6 7 8; do
    /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
    rc=$?
    if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
    sleep 150
done
exit 3
