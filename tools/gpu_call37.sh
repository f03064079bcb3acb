#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_train_kernels.py -q -x 2>&1 | tail -5 > gpurun_out/pytest_37.log
for g in 0 4 2 8; do
DSG_L2_GROUP=$g timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/table_l2g$g.json > gpurun_out/bench_l2g$g.log 2>&1
done
