#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <extra gpurun flags or ""> <command...>
# Retries a gpurun call while the pod answers "busy" (exit code 3: nothing charged).
T=$1; shift
FLAGS=$1; shift
for i in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout $T $FLAGS -- "$@" > /tmp/gpurun_last.log 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then break; fi
    sleep 90
done
tail -80 /tmp/gpurun_last.log
exit $rc
