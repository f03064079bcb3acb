#!/bin/bash
# final-state check of round 1: full GPU suite, smoke, headline bench, launch list of the sampling step
mkdir -p gpurun_out/r1h
O=gpurun_out/r1h
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 --profile-out $O/launch_table_events.json > $O/bench.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/ncu_launches.csv python tools/profile_step.py > $O/ncu_launches_run.log 2>&1
python tools/summarize_launches.py $O/ncu_launches.csv > $O/ncu_launch_summary.txt 2>&1
timeout 600 python bench.py --scheduler ddim --size 512 --batch 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_c4_ddim512.log 2>&1
timeout 900 python bench.py --workload train --steps 8 --warmup 3 --profile-out $O/train_launch_table_events.json > $O/bench_train.log 2>&1
