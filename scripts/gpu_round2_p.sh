#!/bin/bash
# round 2, visit P (1 GPU): parse window 19 / 17 / 15 (prebuilt library variants under variants/; 17 and 15 sort 18-symbol keys = 7 radix passes)
mkdir -p gpurun_out
export MMB_CACHE_DIR=/tmp/mmb_cache
cp mumemto_b200/lib/libmumemto_b200.so /tmp/lib_default.so
for w in 19 17 15; do
  cp variants/libmumemto_b200_w$w.so mumemto_b200/lib/libmumemto_b200.so
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 14 > gpurun_out/p_bench_c2_w$w.json 2> gpurun_out/p_bench_c2_w$w.err; echo "exit $?" >> gpurun_out/p_bench_c2_w$w.err
  timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 14 > gpurun_out/p_bench_c5p_w$w.json 2> gpurun_out/p_bench_c5p_w$w.err; echo "exit $?" >> gpurun_out/p_bench_c5p_w$w.err
  [ -f /tmp/mmb_cache/c4_r0_w1.npy ] || { cp /tmp/mmb_cache/c2_r0_w1.npy /tmp/mmb_cache/c4_r0_w1.npy; cp /tmp/mmb_cache/c2_r0_w1.npy.sizes.npy /tmp/mmb_cache/c4_r0_w1.npy.sizes.npy; }
  timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 14 > gpurun_out/p_bench_c4_w$w.json 2> gpurun_out/p_bench_c4_w$w.err; echo "exit $?" >> gpurun_out/p_bench_c4_w$w.err
done
cp /tmp/lib_default.so mumemto_b200/lib/libmumemto_b200.so
for w in 19 17 15; do for f in c2 c5p c4; do echo == $f w$w; head -c 215 gpurun_out/p_bench_${f}_w$w.json | tail -c 75; tail -n 1 gpurun_out/p_bench_${f}_w$w.err; done; done
