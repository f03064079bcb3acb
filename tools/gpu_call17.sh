#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train_kernels.py -q -x -k "attention_bwd" 2>&1 | tail -30 > gpurun_out/pytest_attn_bwd.log
timeout 600 python -m pytest tests/test_gpu_train_unet.py tests/test_gpu_configs.py tests/test_gpu_attention_tc.py -q 2>&1 | tail -8 > gpurun_out/pytest_new.log
timeout 900 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/train_table.json > gpurun_out/bench_train.log 2>&1
