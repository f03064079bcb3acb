#!/bin/bash
# round-1 evidence refresh (r1g): full GPU test-suite, smoke, both bench arms, training / 512 DDIM configs, launch lists,
# ncu --set full captures of the hot kernels, op sweep.
mkdir -p gpurun_out/r1g
O=gpurun_out/r1g
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $O/bench_reference.log 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 --profile-out $O/launch_table_events.json > $O/bench.log 2>&1
timeout 900 python bench.py --workload train --steps 8 --warmup 3 --profile-out $O/train_launch_table_events.json > $O/bench_train.log 2>&1
timeout 600 python bench.py --scheduler ddim --size 512 --batch 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_c4_ddim512.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/ncu_launches.csv python tools/profile_step.py > $O/ncu_launches_run.log 2>&1
python tools/summarize_launches.py $O/ncu_launches.csv > $O/ncu_launch_summary.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/ncu_train_launches.csv python tools/profile_train_step.py > $O/ncu_train_launches_run.log 2>&1
python tools/summarize_launches.py $O/ncu_train_launches.csv > $O/ncu_train_launch_summary.txt 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:igemm \
    -o /tmp/prof_conv python tools/profile_step.py > $O/ncu_conv.log 2>&1
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > $O/prof_conv_raw.csv 2>/dev/null
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"gn_apply|attention|conv_in|sched|temb" \
    -o /tmp/prof_other python tools/profile_step.py > $O/ncu_other.log 2>&1
ncu -i /tmp/prof_other.ncu-rep --page raw --csv > $O/prof_other_raw.csv 2>/dev/null
timeout 600 python tools/op_sweep.py --raster --out $O/op_sweep.json > $O/op_sweep.log 2>&1
du -sh $O
