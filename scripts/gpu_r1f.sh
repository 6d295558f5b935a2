#!/bin/bash
# Visit 6: parity, A/B of the expand_members variants and of the stream tile size on C2, ncu full of the doubling-round kernel.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_r1f.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 0 1 2 3; do
  echo "=== bench c2 expand variant $v"
  MMB_EXPAND_VARIANT=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/c2_err.log | tee gpurun_out/bench_c2_v$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['stages_ms'], d['matches'])"
done
for t in 536870912 1073741824; do
  echo "=== bench c2 tile $t"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --tile-ranks $t 2>> gpurun_out/c2_err.log | tee gpurun_out/bench_c2_t$t.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['stages_ms'], d['matches'])"
done
tail -3 gpurun_out/c2_err.log
echo "=== ncu full sort_groups_warp"
MMB_MIN_WARMUP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sort_groups_warp -s 30 -c 1 -f -o gpurun_out/prof_c2_sort_groups_warp python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_sgw.log 2>&1
tail -2 gpurun_out/ncu_full_sgw.log | cut -c1-200
echo "=== done"
