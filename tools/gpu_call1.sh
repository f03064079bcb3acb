#!/bin/bash
# round-1 GPU session: parity tests, conv kernel A/B, bench, launch list, ncu captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest.log
python tools/conv_bench.py --out gpurun_out/conv_bench.json > gpurun_out/conv_bench.log 2>&1
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/table.json > gpurun_out/bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
for n in 64 128 256; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:igemm_halo_kernel<\\(int\\)${n}," -s 12 -c 2 -o gpurun_out/prof_halo${n} \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_halo${n}.log 2>&1
done
ls -la gpurun_out
