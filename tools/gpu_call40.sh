#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"igemm_kernel" \
    -o gpurun_out/prof_igemm python tools/profile_step.py > gpurun_out/ncu_igemm.log 2>&1
timeout 300 python tools/op_sweep.py --raster-only --sizes 256,512 --out gpurun_out/raster_sweep.json > gpurun_out/raster_sweep.log 2>&1
