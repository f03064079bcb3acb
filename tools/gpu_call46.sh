#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_unet.py tests/test_gpu_configs.py -q -x 2>&1 | tail -5 > gpurun_out/pytest_46.log
rm -f gpurun_out/bench_co0.log gpurun_out/bench_co1.log
for m in 0 1 0 1; do
DSG_CONV_OUT_MMA=$m timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_co$m.log 2>&1
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"conv_out_mma|conv_in_mma" \
    --log-file gpurun_out/small_launches.csv python tools/profile_step.py > /dev/null 2>&1
