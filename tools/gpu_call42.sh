#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_unet.py tests/test_gpu_configs.py tests/test_gpu_train_unet.py -q -x 2>&1 | tail -8 > gpurun_out/pytest_42.log
for m in 0 1 0 1; do
DSG_CONV_OUT_MMA=$m timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/table_co$m.json >> gpurun_out/bench_co$m.log 2>&1
done
