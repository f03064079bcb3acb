#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -q -x --deselect tests/test_gpu_train_kernels.py::test_conv_wgrad_matches_autograd 2>&1 | tail -40 > gpurun_out/pytest_train.log
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -q -k "wgrad" 2>&1 | tail -60 > gpurun_out/pytest_wgrad.log
