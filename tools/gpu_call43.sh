#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"conv_out_mma|conv_in_mma" \
    -o gpurun_out/prof_small python tools/profile_step.py > gpurun_out/ncu_small.log 2>&1
