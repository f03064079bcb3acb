#!/bin/bash
mkdir -p gpurun_out
for m in 1 2 0; do
DSG_FUSE_GN=$m timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/table_fuse$m.json > gpurun_out/bench_fuse$m.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_configs.py tests/test_gpu_fused_gn_conv.py -q 2>&1 | tail -12 > gpurun_out/pytest_all.log
