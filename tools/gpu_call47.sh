#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_unet.py -q -x 2>&1 | tail -5 > gpurun_out/pytest_47.log
timeout 600 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/train_table.json > gpurun_out/bench_train_47.log 2>&1
