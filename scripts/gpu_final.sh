#!/bin/bash
# Round-end visit: parity, the default bench line (C2, 10 timed steps, end-to-end leg, CPU baseline), the reference arm, -M mode on a C5
# partition, ncu --set full of the two stream kernels on C2, and an ncu launch list of a 64 x 16 Mbp run (a C2 list costs > 15 GPU-minutes).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_final.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== bench c2 (default)"; timeout 900 python bench.py 2> gpurun_out/c2_err.log | tee gpurun_out/bench_c2.json | cut -c1-2600; tail -3 gpurun_out/c2_err.log
echo "=== reference arm"; timeout 300 python bench.py --impl reference --steps 1 --warmup 0 | tee gpurun_out/bench_c2_reference.json | cut -c1-700
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["value"]/1e9, d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["stages_ms"], d["matches"], d["verified_sample_lines"])'
echo "=== bench c5 partition, -M -n, one GPU"; timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/c2_err.log | tee gpurun_out/bench_c5_merge.json | python -c "$show"
for k in expand_members detect_chunk_kernel; do
  echo "=== ncu full $k"
  MMB_MIN_WARMUP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 14 -c 1 -f -o gpurun_out/prof_c2_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_$k.log | cut -c1-200
done
echo "=== ncu launch list h64_16m"
MMB_MIN_WARMUP=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_h64_16m.csv python bench.py --workload h64_16m --route pfp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/launches_h64_16m.csv
echo "=== done"
