#!/bin/bash
# round 2, visit Q (1 GPU): final single-GPU lines of the round (window 15): GPU tests, the default bench line (parity sample, golden md5,
# e2e, CPU baseline, CLI end to end), C3, C4, one C5 partition, 64 x 40 Mbp (text crosses 2^32 symbols; golden md5), ncu launch list and
# --set full captures of the kernels that lead the table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/q_pytest.log
tail -n 4 gpurun_out/q_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python bench.py --steps 10 --warmup 3 --kernel-table-rows 60 --cli-e2e > gpurun_out/q_bench_c2.json 2> gpurun_out/q_bench_c2.err; echo "exit $?" >> gpurun_out/q_bench_c2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 60 > gpurun_out/q_bench_c5p.json 2> gpurun_out/q_bench_c5p.err; echo "exit $?" >> gpurun_out/q_bench_c5p.err
cp /tmp/mmb_cache/c2_r0_w1.npy /tmp/mmb_cache/c4_r0_w1.npy; cp /tmp/mmb_cache/c2_r0_w1.npy.sizes.npy /tmp/mmb_cache/c4_r0_w1.npy.sizes.npy
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 --kernel-table-rows 16 > gpurun_out/q_bench_c4.json 2> gpurun_out/q_bench_c4.err; echo "exit $?" >> gpurun_out/q_bench_c4.err
mv /tmp/mmb_cache/c4_r0_w1.npy /tmp/mmb_cache/c3_r0_w1.npy; mv /tmp/mmb_cache/c4_r0_w1.npy.sizes.npy /tmp/mmb_cache/c3_r0_w1.npy.sizes.npy
timeout 900 python bench.py --workload c3 --steps 5 --warmup 3 --kernel-table-rows 16 > gpurun_out/q_bench_c3.json 2> gpurun_out/q_bench_c3.err; echo "exit $?" >> gpurun_out/q_bench_c3.err
rm -f /tmp/mmb_cache/c3_r0_w1.npy
timeout 900 python bench.py --workload h64_40m --steps 3 --warmup 3 --no-cpu-baseline --kernel-table-rows 8 > gpurun_out/q_bench_h64_40m.json 2> gpurun_out/q_bench_h64_40m.err; echo "exit $?" >> gpurun_out/q_bench_h64_40m.err
for f in c2 c5p c4 c3 h64_40m; do echo == $f; head -c 250 gpurun_out/q_bench_$f.json; tail -n 2 gpurun_out/q_bench_$f.err; done
export MMB_MIN_WARMUP=1
NCU_BENCH="python bench.py --workload h64_16m --route pfp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --kernel-table-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches_h64_16m.csv $NCU_BENCH > gpurun_out/q_ncu_list.log 2>&1
for k in rank_words classify_segments parse_lcp trigger_masks; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/q_ncu_$k $NCU_BENCH > gpurun_out/q_ncu_$k.log 2>&1
done
