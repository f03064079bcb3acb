#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest.log
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 60 -c 3 -o gpurun_out/prof_gn \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -o gpurun_out/prof_attn \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out
