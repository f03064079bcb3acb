#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest.log
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
ls -la gpurun_out
