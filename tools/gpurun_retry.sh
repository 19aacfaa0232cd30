#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (re-submits while the pod answers transient / busy)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|retry in a minute" || [ $rc -eq 3 ]; then
    echo "[retry $i] pod busy, waiting" ; sleep 150; continue
  fi
  echo "$out" | tail -40
  exit $rc
done
echo "gave up"; exit 3
