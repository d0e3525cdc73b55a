#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <logfile> <command...>; retries while the pod answers busy (rc 3)
t=$1; log=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 120
done
