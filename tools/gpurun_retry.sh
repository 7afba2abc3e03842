#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'  - resubmits while the pod answers "transient" / busy
t=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=refused\|status=busy"; then sleep 90; continue; fi
  break
done
