#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "transient/busy" (nothing charged)
T=$1; shift
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$OUT" | tail -80
  if echo "$OUT" | grep -q "status=transient\|retry in a few minutes\|no box\|busy"; then
    if echo "$OUT" | grep -q "charged=0.0s"; then echo "[retry $i] sleeping 150 s"; sleep 150; continue; fi
  fi
  break
done
