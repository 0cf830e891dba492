#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <command...> — retries while the pod answers busy (nothing is charged for those)
T=$1; shift
for i in $(seq 1 30); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up: pod busy"; exit 3
