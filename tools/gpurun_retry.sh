#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'   (retries while the pod answers "transient")
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -70
  if echo "$out" | grep -q "status=transient\|answers busy\|status=busy"; then
    sleep 150
    continue
  fi
  break
done
