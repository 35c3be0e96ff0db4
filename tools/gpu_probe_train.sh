#!/bin/bash
# GPU call (short): the training-step tests + the step timing.  usage: gpurun --timeout 420 -- 'bash tools/gpu_probe_train.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
K='batchnorm or conv_backward or wgrad_tc or tcgen05_vs_fp32 or lstm_layer_backward or training or train_forward'
timeout 330 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$K" > gpurun_out/probe_train.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/probe_train.log
HN_TRAIN_PROF=1 timeout 150 python tools/train_bench.py 8 3 > gpurun_out/train_step.json 2> gpurun_out/train_step.err
echo "step:"; cat gpurun_out/train_step.json
timeout 100 python tools/train_bench.py 8 5 > gpurun_out/train_step_noprof.json 2>> gpurun_out/train_step.err
echo "step (no per-unit events):"; cat gpurun_out/train_step_noprof.json
