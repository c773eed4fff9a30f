#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout-seconds> <log> <command...>   -- retries while the pod answers busy (exit 3)
T=$1; LOG=$2; shift 2
G=""; [ -n "$GPUS" ] && G="--gpus $GPUS"
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
