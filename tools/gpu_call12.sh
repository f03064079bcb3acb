#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_unet.py -q -x 2>&1 | tail -60 > gpurun_out/pytest_train_unet.log
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -q 2>&1 | tail -5 > gpurun_out/pytest_train.log
