#!/bin/bash
# final commit check: full GPU suite, smoke, headline bench
mkdir -p gpurun_out/r1j
O=gpurun_out/r1j
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 > $O/bench.log 2>&1
