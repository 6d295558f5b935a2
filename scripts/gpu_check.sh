#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench lines, ncu launch list, memcheck.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
echo "=== build"; timeout 600 python -c "import __graft_entry__ as g; g.build()" || echo BUILD_FAILED
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
echo "=== memcheck (tiny)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
echo "=== racecheck (tiny)"; timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "=== debug repeats"; MMB_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -s -k repeats 2>&1 | grep -v "^\[prims\] temp storage [0-9]\{1,7\} bytes" | tail -60
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -40
echo "=== bench c1"; timeout 900 python bench.py --steps 5 --warmup 3 | tee gpurun_out/bench_c1.json
echo "=== bench h64_1m"; timeout 900 python bench.py --workload h64_1m --steps 5 --warmup 3 --no-cpu-baseline | tee gpurun_out/bench_h64_1m.json
echo "=== bench h64_4m"; timeout 900 python bench.py --workload h64_4m --steps 3 --warmup 3 --no-cpu-baseline | tee gpurun_out/bench_h64_4m.json
echo "=== ncu launch list (c1, 1 step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_c1.csv
echo "=== ncu launch list (h64_1m, 1 step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_h64_1m.csv python bench.py --workload h64_1m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
tail -2 gpurun_out/ncu_bench2.log | cut -c1-200; wc -l gpurun_out/launches_h64_1m.csv
echo "=== done"
