#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_unet.py tests/test_gpu_configs.py -q 2>&1 | tail -30 > gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_all.log
