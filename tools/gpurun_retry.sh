#!/bin/bash
# Retries a gpurun call while the pod answers "transient" (no slot free; nothing charged).
#   tools/gpurun_retry.sh <log file> <timeout s> '<command>'
log=$1; to=$2; cmd=$3
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$cmd" > "$log" 2>&1
  if grep -q "status=transient\|exit code 3\|rc=3" "$log"; then sleep 45; continue; fi
  break
done
