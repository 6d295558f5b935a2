#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_c2.log) 2>&1
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv; free -g | head -2; nproc
echo "=== bench c2"; timeout 2400 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/c2_err.log | tee gpurun_out/bench_c2.json | cut -c1-1500
tail -20 gpurun_out/c2_err.log
nvidia-smi --query-gpu=memory.used --format=csv
echo "=== done"
