#!/bin/bash
# round 2, visit H (1 GPU): text stage in two passes, staged phrase kernels, warp merges, pipelined upload, MEM pre-check
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/h_pytest.log
tail -n 4 gpurun_out/h_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python bench.py --steps 5 --warmup 3 --kernel-table-rows 40 --cli-e2e > gpurun_out/h_bench_c2.json 2> gpurun_out/h_bench_c2.err; echo "exit $?" >> gpurun_out/h_bench_c2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 40 > gpurun_out/h_bench_c5p.json 2> gpurun_out/h_bench_c5p.err; echo "exit $?" >> gpurun_out/h_bench_c5p.err
cp /tmp/mmb_cache/c2_r0_w1.npy /tmp/mmb_cache/c4_r0_w1.npy; cp /tmp/mmb_cache/c2_r0_w1.npy.sizes.npy /tmp/mmb_cache/c4_r0_w1.npy.sizes.npy
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --kernel-table-rows 12 > gpurun_out/h_bench_c4.json 2> gpurun_out/h_bench_c4.err; echo "exit $?" >> gpurun_out/h_bench_c4.err
for f in c2 c5p c4; do echo == $f; head -c 300 gpurun_out/h_bench_$f.json; tail -n 3 gpurun_out/h_bench_$f.err; done
