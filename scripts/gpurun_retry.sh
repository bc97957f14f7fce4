#!/bin/bash
# usage: gpurun_retry.sh <gpurun args...>   -- retries while the pod has no free slot (transient, nothing charged)
for attempt in $(seq 1 12); do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    echo "$out" | tail -40
    if echo "$out" | grep -q "status=transient"; then echo "[retry] attempt $attempt: no slot, sleeping 150 s"; sleep 150; continue; fi
    exit 0
done
exit 3
