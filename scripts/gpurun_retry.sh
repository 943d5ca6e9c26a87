#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3: nothing charged).  usage: gpurun_retry.sh <gpurun args...>
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
