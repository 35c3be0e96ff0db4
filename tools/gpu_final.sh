#!/bin/bash
# GPU call 2 (final state of the round): the whole GPU suite, the default bench line, ncu evidence for the weight-gradient kernel.
# usage (here): gpurun --timeout 780 -- 'bash tools/gpu_final.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$(date +%s)
timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -4 gpurun_out/final_pytest.log
timeout 330 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))"; head -c 600 gpurun_out/bench_n1.json; echo
timeout 150 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 40 -c 3 -f -o gpurun_out/wgrad_tc_full \
    python tools/train_bench.py 8 1 > gpurun_out/wgrad_ncu.log 2>&1
echo "ncu full rc=$? t=$(( $(date +%s) - T0 ))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_bench.py 8 1 > gpurun_out/train_ncu.log 2>&1
echo "ncu list rc=$? t=$(( $(date +%s) - T0 ))"
HN_WGRAD_TC=${OTHER_WGRAD:-0} timeout 200 python -m pytest tests -q -m gpu -p no:cacheprovider -k "training" > gpurun_out/final_pytest_other.log 2>&1
echo "pytest(other wgrad setting) rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/final_pytest_other.log
