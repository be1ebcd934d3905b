#!/bin/bash
# Retry a gpurun call while the pod answers "transient" (exit 3): bash tools/gpu_retry.sh <logfile> <timeout> <command...>
LOG=$1; TMO=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 120
done
