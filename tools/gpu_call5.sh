#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention_tc.py -q -x 2>&1 | tail -40 > gpurun_out/pytest_attn.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_attention_tc.py 2>&1 | tail -25 > gpurun_out/pytest.log
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
ls -la gpurun_out
