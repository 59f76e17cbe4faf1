#!/bin/bash
# retry.sh <timeout> <command...>: gpurun with retries while the pod answers "busy" (exit 3) -- nothing is charged for those
t=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
