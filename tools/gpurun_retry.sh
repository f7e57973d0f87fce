#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args ...] -- 'command'   (retries while the pod has no free slot)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
