#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_igemm.py -q -x -k "cta_pair" 2>&1 | tail -30 > gpurun_out/pytest_pair.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_igemm.py::test_igemm_halo_cta_pair 2>&1 | tail -25 > gpurun_out/pytest.log
timeout 600 python tools/conv_bench.py --out gpurun_out/conv_bench.json > gpurun_out/conv_bench.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
ls -la gpurun_out
