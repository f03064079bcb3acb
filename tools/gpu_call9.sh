#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest.log
timeout 900 python bench.py --steps 40 --warmup 5 --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/launches_run.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm \
    -o gpurun_out/prof_conv python tools/profile_step.py > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gn_apply|attention|conv_in" \
    -c 12 -o gpurun_out/prof_other python tools/profile_step.py > gpurun_out/ncu_other.log 2>&1
ls -la gpurun_out
