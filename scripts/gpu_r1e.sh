#!/bin/bash
# Visit 5: parity + C2 bench of v7c, A/B of the expand key cache, ncu full of sort_groups_warp<4,1>, expand_members, detect_chunk.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_r1e.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== bench c2"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/c2_err.log | tee gpurun_out/bench_c2.json | cut -c1-1900; tail -3 gpurun_out/c2_err.log
echo "=== bench c2 no key cache"; MMB_EXPAND_CACHE=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/c2_err.log | tee gpurun_out/bench_c2_nocache.json | cut -c1-1500
echo "=== phases"; MMB_MIN_WARMUP=1 MMB_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2> gpurun_out/c2_phases.log | cut -c1-100
grep '^\[phase\]' gpurun_out/c2_phases.log | python -c "
import sys, collections
tot=collections.OrderedDict(); n=collections.Counter()
lines=[l for l in sys.stdin]
for l in lines[len(lines)//2:]:
    name=l[8:37].strip(); ms=float(l.split()[-2]); tot[name]=tot.get(name,0)+ms; n[name]+=1
for k,v in tot.items(): print('%-30s %9.2f ms x%d'%(k,v,n[k]))
"
for k in "sort_groups_warp<4" expand_members detect_chunk_kernel; do
  n=$(echo $k | tr -d '<')
  echo "=== ncu full $k"
  MMB_MIN_WARMUP=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 20 -c 1 -f -o gpurun_out/prof_c2_$n python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$n.log 2>&1
  tail -2 gpurun_out/ncu_full_$n.log | cut -c1-200
done
echo "=== done"
