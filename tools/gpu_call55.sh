#!/bin/bash
# A/B: gn_bwd_apply_kernel at 2 (default build) vs 3 (variant build) resident CTAs per SM, same box
mkdir -p gpurun_out
L=drivescenegen_b200/libdsg_b200.so
cp $L /tmp/lib_a.so
for v in a b a b; do
  if [ $v = b ]; then cp drivescenegen_b200/libdsg_b200_variant.bin $L; else cp /tmp/lib_a.so $L; fi
  timeout 300 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gnb_$v.log 2>&1
  tail -1 gpurun_out/bench_gnb_$v.log >> gpurun_out/bench_gnb_all_$v.log
done
cp /tmp/lib_a.so $L
