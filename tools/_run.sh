mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_unet.py -m gpu -q -x 2>&1 | grep -v "it/s" | tail -12 > $O/pytest_a.log
tail -4 $O/pytest_a.log | cut -c1-300
echo unfused; DSG_GN_BWD_FUSED=0 timeout 300 python tools/gn_bwd_bench.py 2>&1 | grep "GB/s"
echo fused ilp2; timeout 300 python tools/gn_bwd_bench.py 2>&1 | grep "GB/s"
echo fused ilp3; DSG_LIB=$PWD/drivescenegen_b200/libdsg_v1.bin timeout 300 python tools/gn_bwd_bench.py 2>&1 | grep "GB/s"
for v in 0 1 0 1; do DSG_GN_BWD_FUSED=$v timeout 300 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_train_f$v.log 2>&1; tail -1 $O/bench_train_f$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fused=$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(x,2) for k,x in d['breakdown'].items()})"; done
