#!/bin/bash
# GPU call 1 (short): does the tcgen05 weight-gradient kernel produce the right numbers, and what does it buy?
# usage (here): gpurun --timeout 420 -- 'bash tools/gpu_probe_wgrad.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/wgrad_tc_errors.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe_gpu.txt 2>&1
K='wgrad_tc or tcgen05_vs_fp32'
timeout 300 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$K" > gpurun_out/probe_desc0.log 2>&1
rc0=$?
echo "desc0 rc=$rc0"; tail -5 gpurun_out/probe_desc0.log
if [ $rc0 -ne 0 ]; then
  HN_WGRAD_TC_DESC=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$K" > gpurun_out/probe_desc1.log 2>&1
  rc1=$?
  echo "desc1 rc=$rc1"; tail -5 gpurun_out/probe_desc1.log
  [ $rc1 -eq 0 ] && export HN_WGRAD_TC_DESC=1
fi
HN_TRAIN_PROF=1 HN_WGRAD_TC=0 timeout 150 python tools/train_bench.py 8 3 > gpurun_out/train_wgrad_off.json 2> gpurun_out/train_wgrad_off.err
echo "off:"; cat gpurun_out/train_wgrad_off.json
HN_TRAIN_PROF=1 HN_WGRAD_TC=1 timeout 150 python tools/train_bench.py 8 3 > gpurun_out/train_wgrad_on.json 2> gpurun_out/train_wgrad_on.err
echo "on:"; cat gpurun_out/train_wgrad_on.json; tail -3 gpurun_out/train_wgrad_on.err
cat gpurun_out/wgrad_tc_errors.jsonl 2>/dev/null | tail -30
