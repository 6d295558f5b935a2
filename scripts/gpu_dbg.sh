#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_dbg.log) 2>&1
echo "=== host profile h64_4m pfp"; MMB_PROFILE_HOST=1 timeout 600 python bench.py --workload h64_4m --route pfp --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | cut -c1-2500
echo "=== host profile h64_4m direct"; MMB_PROFILE_HOST=1 timeout 600 python bench.py --workload h64_4m --route direct --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | cut -c1-1500
echo "=== done"
