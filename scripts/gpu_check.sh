#!/bin/bash
# One GPU-box visit: smoke, sanitizers, GPU parity tests, bench lines, ncu launch lists.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
echo "=== build"; timeout 600 python -c "import __graft_entry__ as g; g.build()" || echo BUILD_FAILED
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -20
echo "=== bench c1"; timeout 900 python bench.py --workload c1 --steps 10 --warmup 3 | tee gpurun_out/bench_c1.json | cut -c1-400
echo "=== bench c2 (default)"; timeout 2400 python bench.py | tee gpurun_out/bench_c2.json | cut -c1-3000
echo "=== bench h64_16m pfp"; timeout 900 python bench.py --workload h64_16m --route pfp --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | cut -c1-1500
echo "=== done"
