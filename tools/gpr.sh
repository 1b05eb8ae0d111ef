#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged).  usage: tools/gpr.sh <timeout s> [--gpus N] -- '<command>'
TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" > /tmp/gpr_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then cat /tmp/gpr_last.log | tail -60; exit $rc; fi
  sleep 90
done
echo "gave up: no GPU slot"; exit 3
