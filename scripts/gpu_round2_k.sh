#!/bin/bash
# round 2, visit K (1 GPU): phrase hash / verification on 16-byte chunks; default bench line with the per-kernel ncu traffic columns
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/k_pytest.log
tail -n 4 gpurun_out/k_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python bench.py --steps 5 --warmup 3 --kernel-table-rows 60 --no-cpu-baseline > gpurun_out/k_bench_c2.json 2> gpurun_out/k_bench_c2.err; echo "exit $?" >> gpurun_out/k_bench_c2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 60 > gpurun_out/k_bench_c5p.json 2> gpurun_out/k_bench_c5p.err; echo "exit $?" >> gpurun_out/k_bench_c5p.err
for f in c2 c5p; do echo == $f; head -c 250 gpurun_out/k_bench_$f.json; tail -n 2 gpurun_out/k_bench_$f.err; done
export MMB_MIN_WARMUP=1
NCU_BENCH="python bench.py --workload h64_16m --route pfp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --kernel-table-steps 0"
for k in phrase_hash verify_groups; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/k_ncu_$k $NCU_BENCH > gpurun_out/k_ncu_$k.log 2>&1
done
