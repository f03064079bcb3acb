#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 40 --warmup 5 --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/launches_run.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:igemm \
    -o /tmp/prof_conv python tools/profile_step.py > gpurun_out/ncu_conv.log 2>&1
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"gn_apply|gn_stats|attention|conv_in|sched|temb" \
    -o /tmp/prof_other python tools/profile_step.py > gpurun_out/ncu_other.log 2>&1
ncu -i /tmp/prof_other.ncu-rep --page raw --csv > gpurun_out/prof_other_raw.csv 2>/dev/null
# one small report with source for the three tile shapes of the CTA-pair conv kernel (level 0 / 1 / 3)
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_halo -s 4 -c 1 \
    -o gpurun_out/prof_halo_l0 python tools/profile_step.py > /dev/null 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_halo -s 22 -c 1 \
    -o gpurun_out/prof_halo_l3 python tools/profile_step.py > /dev/null 2>&1
du -sh gpurun_out; ls -la gpurun_out
