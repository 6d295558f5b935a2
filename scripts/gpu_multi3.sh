#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the default bench line at N ranks (C5 partition shape, exchange + merge inside the step) + reference arm.
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_multi.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/multi_err.log | tee gpurun_out/bench_c5_n${N}.json | cut -c1-3200
tail -3 gpurun_out/multi_err.log
echo "=== done"
