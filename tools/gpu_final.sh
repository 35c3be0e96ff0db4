#!/bin/bash
# Final GPU call of the round: the whole GPU suite, the default bench line, one ncu --set full capture of the weight-gradient kernel.
# usage (here): gpurun --timeout 560 -- 'bash tools/gpu_final.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -4 gpurun_out/final_pytest.log
timeout 240 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))"; head -c 400 gpurun_out/bench_n1.json; echo
# launches of the third step (75 per step, order in tools/ncu_wgrad.py): #11 = ghc_lst.3.layer.0 (9 taps, 2048 -> 1024, 2048 pixels), #22 = ghc_lst.0.layer.0 (256 -> 128, 131072 pixels)
timeout 120 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 161 -c 1 -f -o gpurun_out/wgrad_tc_ghc3_0 \
    python tools/train_bench.py 8 1 > gpurun_out/wgrad_ncu_full.log 2>&1
echo "ncu full (ghc3.0) rc=$? t=$(( $(date +%s) - T0 ))"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 172 -c 1 -f -o gpurun_out/wgrad_tc_ghc0_0 \
    python tools/train_bench.py 8 1 >> gpurun_out/wgrad_ncu_full.log 2>&1
echo "ncu full (ghc0.0) rc=$? t=$(( $(date +%s) - T0 ))"
