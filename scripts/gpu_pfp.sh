#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_pfp.log) 2>&1
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" || echo SMOKE_FAILED
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30
for w in h64_1m h64_4m; do
  echo "=== bench $w direct"; timeout 900 python bench.py --workload $w --route direct --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | tee gpurun_out/bench_${w}_direct.json | cut -c1-600
  echo "=== bench $w pfp"; timeout 900 python bench.py --workload $w --route pfp --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | tee gpurun_out/bench_${w}_pfp.json | cut -c1-600
done
echo "=== bench h64_16m pfp"; timeout 900 python bench.py --workload h64_16m --route pfp --steps 2 --warmup 3 --no-cpu-baseline --no-e2e | tee gpurun_out/bench_h64_16m_pfp.json | cut -c1-600
echo "=== done"
