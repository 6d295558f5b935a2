#!/bin/bash
# round 2, visit G (1 GPU): quick check of the phrase-level parse_lcp: tests + C2 + one C5 partition
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest.log
tail -n 4 gpurun_out/g_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python bench.py --steps 5 --warmup 3 --kernel-table-rows 30 > gpurun_out/g_bench_c2.json 2> gpurun_out/g_bench_c2.err; echo "exit $?" >> gpurun_out/g_bench_c2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 30 > gpurun_out/g_bench_c5p.json 2> gpurun_out/g_bench_c5p.err; echo "exit $?" >> gpurun_out/g_bench_c5p.err
for f in c2 c5p; do echo == $f; head -c 300 gpurun_out/g_bench_$f.json; tail -n 3 gpurun_out/g_bench_$f.err; done
