#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
DSG_PDL=0 timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nopdl.log 2>&1
timeout 900 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/train_table.json > gpurun_out/bench_train.log 2>&1
DSG_PDL=0 timeout 900 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_nopdl.log 2>&1
