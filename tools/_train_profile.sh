mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py 8 1 > gpurun_out/train_ncu.log 2>&1
tail -2 gpurun_out/train_ncu.log
python tools/launch_summary.py gpurun_out/train_launches.csv gpurun_out/train_launches_summary.json "ncu --metrics gpu__time_duration.sum --clock-control none python tools/train_bench.py 8 1" 2>&1 | tail -16
