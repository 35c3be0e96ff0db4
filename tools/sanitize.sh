#!/bin/bash
# compute-sanitizer passes over every kernel of the library (the reference has no
# race detection; this is ours).  Output: gpurun_out/sanitize_<tool>.log; summary lines at the end.
# Usage (GPU box): tools/sanitize.sh [bs]
bs=${1:-2}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_once.py $bs all \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  grep -E 'model ok|aux ok' gpurun_out/sanitize_$tool.log | tr '\n' ' '; echo
done
