#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): partition-merge path, one rank per GPU over NCCL; a small workload first, then the C5 shape.
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_multi.log) 2>&1
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader
for w in h64_4m c2; do
  echo "=== bench $w N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 3 --warmup 3 2> gpurun_out/multi_${w}_err.log | tee gpurun_out/bench_${w}_n${N}.json | cut -c1-3000
  tail -5 gpurun_out/multi_${w}_err.log
done
echo "=== reference arm N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-600
echo "=== done"
