#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/op_sweep.py --out gpurun_out/op_sweep.json > gpurun_out/op_sweep.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/train_launches.csv python tools/profile_train_step.py > gpurun_out/train_launches_run.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"wgrad_kernel" -c 30 \
    -o /tmp/prof_wgrad python tools/profile_train_step.py > gpurun_out/ncu_wgrad.log 2>&1
ncu -i /tmp/prof_wgrad.ncu-rep --page raw --csv > gpurun_out/prof_wgrad_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"gn_bwd_stats|gn_bwd_apply|attention_bwd_tc" -c 25 \
    -o /tmp/prof_gnbwd python tools/profile_train_step.py > gpurun_out/ncu_gnbwd.log 2>&1
ncu -i /tmp/prof_gnbwd.ncu-rep --page raw --csv > gpurun_out/prof_gnbwd_raw.csv 2>/dev/null
du -sh gpurun_out
