#!/bin/bash
# gpurun with retry while the pod is busy (exit 3 = nothing charged).  usage: tools/gpu_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
