#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
DSG_FUSE_GN=0 timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nofuse.log 2>&1
