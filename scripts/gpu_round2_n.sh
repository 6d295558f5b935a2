#!/bin/bash
# round 2, visit N (1 GPU): initial-sort key of 18 / 16 symbols instead of 21 (one or two radix passes fewer, coarser segments)
mkdir -p gpurun_out
export MMB_CACHE_DIR=/tmp/mmb_cache
for ks in 21 18 16; do
  MMB_KEY_SYMBOLS=$ks timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 12 > gpurun_out/n_bench_c2_ks$ks.json 2> gpurun_out/n_bench_c2_ks$ks.err; echo "exit $?" >> gpurun_out/n_bench_c2_ks$ks.err
done
cp /tmp/mmb_cache/c2_r0_w1.npy /tmp/mmb_cache/c4_r0_w1.npy; cp /tmp/mmb_cache/c2_r0_w1.npy.sizes.npy /tmp/mmb_cache/c4_r0_w1.npy.sizes.npy
for ks in 18 16; do
  MMB_KEY_SYMBOLS=$ks timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 12 > gpurun_out/n_bench_c4_ks$ks.json 2> gpurun_out/n_bench_c4_ks$ks.err; echo "exit $?" >> gpurun_out/n_bench_c4_ks$ks.err
done
for f in c2_ks21 c2_ks18 c2_ks16 c4_ks18 c4_ks16; do echo == $f; head -c 230 gpurun_out/n_bench_$f.json; tail -n 1 gpurun_out/n_bench_$f.err; done
