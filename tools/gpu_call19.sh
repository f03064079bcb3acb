#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_unet.py tests/test_gpu_elementwise.py tests/test_gpu_unet.py -q -x 2>&1 | tail -8 > gpurun_out/pytest_all.log
timeout 900 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/train_table.json > gpurun_out/bench_train.log 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
