#!/bin/bash
# usage: tools/gpurun_retry.sh TIMEOUT 'command'   -- retries while the pod answers busy (exit 3), up to 40 minutes
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
