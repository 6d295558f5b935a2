#!/bin/bash
# round 2, visit A: GPU tests, C2 bench line (parity sample + kernel table), one C5 partition in -M -n mode
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; free -g >> gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err; echo "exit $?" >> gpurun_out/a_bench_c2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/a_bench_c5p.json 2> gpurun_out/a_bench_c5p.err; echo "exit $?" >> gpurun_out/a_bench_c5p.err
tail -3 gpurun_out/a_pytest.log; head -c 600 gpurun_out/a_bench_c2.json; tail -5 gpurun_out/a_bench_c2.err
