#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT 'command' -- retries while the pod answers busy (exit code 3)
T=$1; shift
for k in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
