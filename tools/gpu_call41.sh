#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_unet.py tests/test_gpu_train_unet.py tests/test_gpu_fused_gn_conv.py -q -x 2>&1 | tail -4 > gpurun_out/pytest_41.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/table.json > gpurun_out/bench_41.log 2>&1
timeout 600 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_41.log 2>&1
