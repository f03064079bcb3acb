#!/bin/bash
# The one GPU-side runner: `gpurun --timeout T -- 'bash tools/gpu.sh <tag> <stage> [<stage> ...]'`.
# Every stage writes under gpurun_out/<tag>/ (merged back by gpurun); copy what should be judged into profiles/.
#   test [pytest args]   full `-m gpu` suite (or the given pytest selection), tail in pytest_gpu.log
#   smoke                __graft_entry__.smoke()
#   bench [args]         headline bench (sampling C2 + the train sub-record) with the per-launch event table
#   ref                  bench.py --impl reference
#   train [args]         bench.py --workload train
#   c4 [args]            512x512 DDIM, B=8
#   launches             ncu launch list (gpu__time_duration) of one sampling step + summary
#   launches-train       same for one training step
#   full <regex> <name>  ncu --set full of the kernels matching <regex> in the sampling step -> <name>.ncu-rep + raw csv
#   full-train <regex> <name>   same inside the training step
#                        (NCU_EXTRA="-c 12" in the environment limits the number of profiled launches)
#   sweep                tools/op_sweep.py --raster
#   traffic <full-stage name> <B> <S> <commit>   conv DRAM traffic json for bench.py's roofline.traffic
#   parity               full -m gpu suite with -s, keeping the [parity] tolerance lines
# Stages are separated by `--`:  bash tools/gpu.sh r2a test -- bench --steps 40 -- launches
TAG=$1; shift
O=gpurun_out/$TAG
mkdir -p $O
NCU_T="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_F="--profile-from-start off --set full --clock-control none --import-source on"
run_stage() {
  local s=$1; shift
  case $s in
    test)
      if [ $# -eq 0 ]; then set -- tests; fi
      timeout 1800 python -m pytest "$@" -m gpu -q -x 2>&1 | tail -15 > $O/pytest_gpu.log ;;
    smoke) timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1 ;;
    bench) timeout 1200 python bench.py --profile-out $O/launch_table_events.json "$@" > $O/bench.log 2>&1 ;;
    ref) timeout 600 python bench.py --impl reference --steps 4 --warmup 1 "$@" > $O/bench_reference.log 2>&1 ;;
    train) timeout 900 python bench.py --workload train --profile-out $O/train_launch_table_events.json "$@" > $O/bench_train.log 2>&1 ;;
    c4) timeout 600 python bench.py --scheduler ddim --size 512 --batch 8 --steps 20 --warmup 3 --no-cpu-baseline "$@" > $O/bench_c4_ddim512.log 2>&1 ;;
    launches)
      timeout 600 ncu $NCU_T --log-file $O/ncu_launches.csv python tools/profile_step.py "$@" > $O/ncu_launches_run.log 2>&1
      python tools/summarize_launches.py $O/ncu_launches.csv > $O/ncu_launch_summary.txt 2>&1 ;;
    launches-train)
      timeout 600 ncu $NCU_T --log-file $O/ncu_train_launches.csv python tools/profile_train_step.py "$@" > $O/ncu_train_launches_run.log 2>&1
      python tools/summarize_launches.py $O/ncu_train_launches.csv > $O/ncu_train_launch_summary.txt 2>&1 ;;
    full|full-train)
      local script=tools/profile_step.py; [ $s = full-train ] && script=tools/profile_train_step.py
      local rx=$1 name=$2; shift 2
      timeout 1500 ncu $NCU_F $NCU_EXTRA -k regex:"$rx" -o $O/$name python $script "$@" > $O/ncu_$name.log 2>&1
      ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
      python tools/ncu_raw_summary.py $O/${name}_raw.csv > $O/${name}_summary.txt 2>&1 ;;
    sweep) timeout 600 python tools/op_sweep.py --raster --out $O/op_sweep.json > $O/op_sweep.log 2>&1 ;;
    traffic)   # traffic <name of a finished `full` stage> <batch> <size> <commit>: profiles/conv_traffic_b<B>_<S>.json
      python tools/conv_traffic.py $O/${1}_raw.csv $2 $3 $4 $O/conv_traffic_b${2}_${3}.json > $O/conv_traffic_${2}_${3}.log 2>&1 ;;
    parity)    # the tolerance lines the parity tests print
      timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -a "\[parity\]\|reference scripts\]\| passed\| failed" > $O/parity_lines.txt ;;
    *) echo "unknown stage $s" >&2 ;;
  esac
}
args=()
for a in "$@" --; do
  if [ "$a" = "--" ]; then
    [ ${#args[@]} -gt 0 ] && run_stage "${args[@]}"
    args=()
  else
    args+=("$a")
  fi
done
du -sh $O
tail -3 $O/pytest_gpu.log 2>/dev/null
tail -c 1500 $O/bench.log 2>/dev/null
