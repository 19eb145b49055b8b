#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit code 3 / status transient); usage: tools/gpurun_retry.sh <gpurun args...>
for attempt in $(seq 1 12); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if grep -q '"status": "transient"' /root/repo/gpurun_out/.last_call.json 2>/dev/null; then sleep 45; continue; fi
    exit $rc
done
exit 3
