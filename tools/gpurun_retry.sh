#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout> <command>   -- retries while the pod has no free slot (exit code 3 / transient)
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
tail -45 /tmp/gpurun_last.txt
exit $rc
