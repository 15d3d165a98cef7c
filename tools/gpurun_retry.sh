#!/bin/bash
# gpurun with retries while the pod answers "transient/busy" (exit 3 or status=transient): tools/gpurun_retry.sh LOG TIMEOUT 'cmd'
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|busy\|no box" $log && ! grep -q "status=ok" $log; then sleep 90; continue; fi
  break
done
echo "rc=$rc" >> $log
