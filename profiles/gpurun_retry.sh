#!/bin/bash
# usage: [GPUS=2] profiles/gpurun_retry.sh TIMEOUT 'command'   -- retries while the pod answers "busy" (exit 3), nothing is charged for those
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
