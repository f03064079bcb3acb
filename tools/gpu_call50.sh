#!/bin/bash
# last call of round 1: final-state suite + smoke + headline bench, and full ncu reports (with source) of the conv kernel
mkdir -p gpurun_out/r1i
O=gpurun_out/r1i
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 --profile-out $O/launch_table_events.json > $O/bench.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"igemm_halo_kernel" -s 0 -c 1 \
    -o $O/halo64_pair_level0 python tools/profile_step.py > $O/ncu_l0.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"igemm_halo_kernel" -s 20 -c 1 \
    -o $O/halo256_pair_level3 python tools/profile_step.py > $O/ncu_l3.log 2>&1
