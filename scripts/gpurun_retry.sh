#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): gpurun_retry.sh <timeout_s> <command...>
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/gpurun_last.log 2>&1
  if grep -q "status=transient" /tmp/gpurun_last.log; then echo "[retry $i] transient"; sleep 60; continue; fi
  break
done
cat /tmp/gpurun_last.log
