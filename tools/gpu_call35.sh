#!/bin/bash
# 8-GPU validation: the driver's launch line for N = 8 (sampling = independent replicas, training = one all-reduce per step)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/n8_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/bench_n8.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --workload train --steps 8 --warmup 3 > gpurun_out/bench_train_n8.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_samebox.log 2>&1
