#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_unet.py -q -x 2>&1 | tail -40 > gpurun_out/pytest_train_unet.log
timeout 900 python bench.py --workload train --steps 5 --warmup 3 --profile-out gpurun_out/train_table.json > gpurun_out/bench_train.log 2>&1
