#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3, nothing charged).  usage: tools/gpurun_retry.sh <gpurun args...>
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  echo "[retry $i] pod busy, sleeping 150 s"; sleep 150
done
exit 3
