#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_gn_conv.py -q -x 2>&1 | tail -30 > gpurun_out/pytest_fused.log
