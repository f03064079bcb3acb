#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/sanitize_new_kernels.py > gpurun_out/sanitize_plain.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_new_kernels.py > gpurun_out/sanitize_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_new_kernels.py > gpurun_out/sanitize_racecheck.log 2>&1
