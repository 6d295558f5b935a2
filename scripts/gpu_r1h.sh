#!/bin/bash
# Visit 9: parity, C2 bench, -M mode on one C5 partition.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_r1h.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["value"]/1e9, d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["stages_ms"], d["matches"], d["verified_sample_lines"])'
echo "=== bench c2"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>> gpurun_out/err.log | tee gpurun_out/bench_c2.json | python -c "$show"
echo "=== phases"; MMB_MIN_WARMUP=1 MMB_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2> gpurun_out/c2_phases.log | cut -c1-100
grep '^\[phase\]' gpurun_out/c2_phases.log | python -c "
import sys, collections
tot=collections.OrderedDict(); n=collections.Counter()
lines=[l for l in sys.stdin]
for l in lines[len(lines)//2:]:
    name=l[8:37].strip(); ms=float(l.split()[-2]); tot[name]=tot.get(name,0)+ms; n[name]+=1
for k,v in tot.items(): print('%-30s %9.2f ms x%d'%(k,v,n[k]))
"
echo "=== bench c5 partition, -M -n, one GPU"; timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/err.log | tee gpurun_out/bench_c5_merge.json | python -c "$show"
tail -3 gpurun_out/err.log
echo "=== done"
