mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_unet.py tests/test_gpu_train_unet.py tests/test_gpu_attention_tc.py -m gpu -q -x 2>&1 | grep -v "it/s" | tail -8 > $O/pytest_a.log
tail -4 $O/pytest_a.log | cut -c1-250
for v in 0 1 0 1; do DSG_HALO_1X1=$v timeout 300 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline --profile-out $O/train_h$v.json > $O/bench_train_h$v.log 2>&1; tail -1 $O/bench_train_h$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('halo1x1=$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(x,2) for k,x in d['breakdown'].items()})"; done
for v in 0 1; do DSG_HALO_1X1=$v timeout 300 python bench.py --steps 30 --no-train --no-cpu-baseline > $O/bench_h$v.log 2>&1; tail -1 $O/bench_h$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('halo1x1=$v', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['kernels']['conv']['ms'],3))"; done
