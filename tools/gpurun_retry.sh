#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod reports "busy" (exit 3 / transient)
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 14); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); rc=$?; fi
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
