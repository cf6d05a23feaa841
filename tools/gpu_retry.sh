#!/bin/bash
# retry a gpurun call until a GPU slot is free (exit code 3 = no slot, nothing charged)
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'  [extra gpurun flags via GPURUN_FLAGS]
T=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout $T -- "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
