#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_unet.py -q -x 2>&1 | tail -4 > gpurun_out/pytest_32.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/table.json > gpurun_out/bench_32.log 2>&1
