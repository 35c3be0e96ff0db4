#!/bin/bash
# ncu evidence for the training step: per-launch table of the tcgen05 weight-gradient kernel (one batch-8 backward) and
# the launch list of the whole step.  usage (here): gpurun --timeout 400 -- 'bash tools/gpu_profile_train.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__grid_size
timeout 170 ncu --clock-control none --metrics $M -k regex:wgrad_tc_kernel -s 150 -c 75 --csv --log-file gpurun_out/wgrad_launches.csv \
    python tools/train_bench.py 8 1 > gpurun_out/wgrad_ncu.log 2>&1
echo "per-launch rc=$?"; grep -c wgrad_tc_kernel gpurun_out/wgrad_launches.csv
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_bench.py 8 1 > gpurun_out/train_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/train_launches.csv
