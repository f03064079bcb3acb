#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"conv_out_mma" \
    -o gpurun_out/prof_small2 python tools/profile_step.py > gpurun_out/ncu_small2.log 2>&1
