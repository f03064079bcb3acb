#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_unet.py tests/test_gpu_configs.py -q -x 2>&1 | tail -6 > gpurun_out/pytest_34.log
timeout 300 python tools/op_sweep.py --sizes 256,512 --batches 8,16 --out gpurun_out/op_sweep_34.json > gpurun_out/op_sweep_34.log 2>&1
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/table.json > gpurun_out/bench_34.log 2>&1
