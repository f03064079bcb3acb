#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_raster.py tests/test_gpu_configs.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_raster.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pipe.log 2>&1
timeout 300 python tools/op_sweep.py --raster-only --sizes 256,512 --out gpurun_out/raster_sweep.json > gpurun_out/raster_sweep.log 2>&1
