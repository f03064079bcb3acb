mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
for v in 0 1 0 1; do DSG_TMA_OUT=$v timeout 300 python bench.py --workload train --steps 8 --warmup 3 --no-cpu-baseline --profile-out $O/train_tma$v.json > $O/bench_train_tma$v.log 2>&1; tail -1 $O/bench_train_tma$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tma=$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(x,2) for k,x in d['breakdown'].items()})"; done
