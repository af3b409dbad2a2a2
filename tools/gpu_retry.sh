#!/bin/bash
# gpu_retry.sh <timeout> <command...>: gpurun with retries while the pod answers busy (exit code 3)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
