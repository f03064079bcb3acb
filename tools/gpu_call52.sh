#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_unet.py tests/test_gpu_configs.py tests/test_gpu_train_unet.py tests/test_gpu_train_kernels.py -q -x 2>&1 | tail -5 > gpurun_out/pytest_52.log
timeout 600 python tools/stats_cost.py > gpurun_out/stats_cost3.log 2>&1
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_52.log 2>&1
timeout 600 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_52.log 2>&1
