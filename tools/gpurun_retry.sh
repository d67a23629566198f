#!/bin/bash
# retries a gpurun call while the pod answers "busy" (exit code 3); usage: tools/gpurun_retry.sh <timeout> '<command>'
to=$1; shift
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$to" -- "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 60
done
exit 3
