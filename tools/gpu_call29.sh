#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_unet.py -q 2>&1 | tail -4 > gpurun_out/pytest_29.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/train_launches.csv python tools/profile_train_step.py > gpurun_out/train_launches_run.log 2>&1
python tools/summarize_launches.py gpurun_out/train_launches.csv > gpurun_out/train_launch_summary.txt 2>&1
timeout 900 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/train_table.json > gpurun_out/bench_train.log 2>&1
