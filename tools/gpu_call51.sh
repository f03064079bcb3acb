#!/bin/bash
mkdir -p gpurun_out/r1i
O=gpurun_out/r1i
timeout 900 python bench.py --steps 40 --warmup 5 --profile-out $O/launch_table_events.json > $O/bench.log 2>&1
timeout 900 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_train.log 2>&1
