#!/bin/bash
# Visit 4: parity + C2 bench of v7b (chunk detector with staged BWT, one-pass trigger scan, cheaper stream set-up) + ncu full of the detector.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_r1d.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== bench c2"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/c2_err.log | tee gpurun_out/bench_c2.json | cut -c1-1900; tail -3 gpurun_out/c2_err.log
echo "=== phases"; MMB_MIN_WARMUP=1 MMB_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2> gpurun_out/c2_phases.log | cut -c1-100
grep '^\[phase\]' gpurun_out/c2_phases.log | python -c "
import sys, collections
tot=collections.OrderedDict(); n=collections.Counter()
lines=[l for l in sys.stdin]
for l in lines[len(lines)//2:]:
    name=l[8:37].strip(); ms=float(l.split()[-2]); tot[name]=tot.get(name,0)+ms; n[name]+=1
for k,v in tot.items(): print('%-30s %9.2f ms x%d'%(k,v,n[k]))
"
for k in detect_chunk_kernel; do
  echo "=== ncu full $k"
  MMB_MIN_WARMUP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -f -o gpurun_out/prof_c2_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_$k.log | cut -c1-200
done
echo "=== done"
