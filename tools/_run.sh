mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_unet.py -m gpu -q -x 2>&1 | grep -v "it/s" | tail -12 > $O/pytest_a.log
tail -5 $O/pytest_a.log | cut -c1-300
for v in 1 2; do timeout 300 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline --profile-out $O/train_$v.json > $O/bench_train_$v.log 2>&1; tail -1 $O/bench_train_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(x,2) for k,x in d['breakdown'].items()}, round(d['roofline']['frac'],3))"; done
