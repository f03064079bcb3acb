#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_train_kernels.py tests/test_gpu_train_unet.py -q -x 2>&1 | tail -4 > gpurun_out/pytest_31.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/train_launches.csv python tools/profile_train_step.py > gpurun_out/train_launches_run.log 2>&1
python tools/summarize_launches.py gpurun_out/train_launches.csv > gpurun_out/train_launch_summary.txt 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attention_tc|conv_in_mma" \
    -o gpurun_out/prof_attn python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1
