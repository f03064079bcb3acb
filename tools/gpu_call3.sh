#!/bin/bash
mkdir -p gpurun_out
./tools/umma_probe > gpurun_out/umma_probe.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest.log
python tools/conv_bench.py --out gpurun_out/conv_bench.json > gpurun_out/conv_bench.log 2>&1
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
ls -la gpurun_out
