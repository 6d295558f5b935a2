#!/bin/bash
# round 2, visit O (1 GPU): window w = 19 (every valid dictionary suffix has >= 20 symbols), pair-wise bitonic stages, narrower large-group sort
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/o_pytest.log
tail -n 4 gpurun_out/o_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python bench.py --steps 5 --warmup 3 --kernel-table-rows 60 --no-cpu-baseline > gpurun_out/o_bench_c2.json 2> gpurun_out/o_bench_c2.err; echo "exit $?" >> gpurun_out/o_bench_c2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 60 > gpurun_out/o_bench_c5p.json 2> gpurun_out/o_bench_c5p.err; echo "exit $?" >> gpurun_out/o_bench_c5p.err
cp /tmp/mmb_cache/c2_r0_w1.npy /tmp/mmb_cache/c4_r0_w1.npy; cp /tmp/mmb_cache/c2_r0_w1.npy.sizes.npy /tmp/mmb_cache/c4_r0_w1.npy.sizes.npy
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 16 > gpurun_out/o_bench_c4.json 2> gpurun_out/o_bench_c4.err; echo "exit $?" >> gpurun_out/o_bench_c4.err
for f in c2 c5p c4; do echo == $f; head -c 250 gpurun_out/o_bench_$f.json; tail -n 2 gpurun_out/o_bench_$f.err; done
