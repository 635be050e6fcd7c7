#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> [--gpus N] -- <command>     retries while the pod answers "busy" (rc 3)
t="$1"; shift
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout "$t" "$@" > /tmp/gpurun_retry.$$ 2>&1
    rc=$?
    if grep -q "status=transient" /tmp/gpurun_retry.$$ || [ $rc -eq 3 ]; then
        echo "[retry $attempt] busy" ; sleep 120 ; continue
    fi
    cat /tmp/gpurun_retry.$$ | tail -40
    exit $rc
done
echo "gave up"; exit 3
