#!/bin/bash
# round 2, visit E: after the per-rank kernel work (segment classes by one walk per segment, thread-per-rank thresholds, Kasai chunks of 256)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -n 4 gpurun_out/e_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python bench.py --steps 5 --warmup 3 --cli-e2e --kernel-table-rows 40 > gpurun_out/e_bench_c2.json 2> gpurun_out/e_bench_c2.err; echo "exit $?" >> gpurun_out/e_bench_c2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 40 > gpurun_out/e_bench_c5p.json 2> gpurun_out/e_bench_c5p.err; echo "exit $?" >> gpurun_out/e_bench_c5p.err
export MMB_MIN_WARMUP=1
NCU_BENCH="python bench.py --workload h64_16m --route pfp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --kernel-table-steps 0"
for k in classify_segments parse_lcp; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/e_ncu_$k $NCU_BENCH > gpurun_out/e_ncu_$k.log 2>&1
done
NCU_M="python bench.py --workload h64_16m --route pfp --merge-mode --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --kernel-table-steps 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:segment_thresholds_masks -c 1 -o gpurun_out/e_ncu_segment_thresholds_masks $NCU_M > gpurun_out/e_ncu_segthr.log 2>&1
for f in c2 c5p; do echo == $f; head -c 300 gpurun_out/e_bench_$f.json; tail -n 3 gpurun_out/e_bench_$f.err; done
