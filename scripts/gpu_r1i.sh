#!/bin/bash
# Visit: parity, -M mode on one C5 partition, C2 (3 steps each) after the cheaper candidate verification.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_r1i.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["value"]/1e9, d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["stages_ms"], d["matches"], d["verified_sample_lines"])'
echo "=== bench c5 partition, -M -n, one GPU"; timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/err.log | tee gpurun_out/bench_c5_merge.json | python -c "$show"
echo "=== bench c2"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>> gpurun_out/err.log | tee gpurun_out/bench_c2.json | python -c "$show"
tail -3 gpurun_out/err.log
echo "=== done"
