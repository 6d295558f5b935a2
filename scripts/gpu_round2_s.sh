#!/bin/bash
# round 2, visit S, T (1 GPU): classify_segments through warp prefix sums
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s_pytest.log
tail -n 3 gpurun_out/s_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python bench.py --steps 10 --warmup 3 --kernel-table-rows 60 --cli-e2e > gpurun_out/s_bench_c2.json 2> gpurun_out/s_bench_c2.err; echo "exit $?" >> gpurun_out/s_bench_c2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 30 > gpurun_out/s_bench_c5p.json 2> gpurun_out/s_bench_c5p.err; echo "exit $?" >> gpurun_out/s_bench_c5p.err
cp /tmp/mmb_cache/c2_r0_w1.npy /tmp/mmb_cache/c4_r0_w1.npy; cp /tmp/mmb_cache/c2_r0_w1.npy.sizes.npy /tmp/mmb_cache/c4_r0_w1.npy.sizes.npy
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 12 > gpurun_out/s_bench_c4.json 2> gpurun_out/s_bench_c4.err; echo "exit $?" >> gpurun_out/s_bench_c4.err
for f in c2 c5p c4; do echo == $f; head -c 250 gpurun_out/s_bench_$f.json; tail -n 2 gpurun_out/s_bench_$f.err; done
export MMB_MIN_WARMUP=1
NCU_BENCH="python bench.py --workload h64_16m --route pfp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --kernel-table-steps 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:classify_segments -c 1 -o gpurun_out/s_ncu_classify_segments $NCU_BENCH > gpurun_out/s_ncu_classify.log 2>&1
