#!/bin/bash
# Visit 2 of the re-entry: parity with the chunked detector, C2 bench A/B (chunk vs window detector), phases, ncu launch list of one warm
# C2 step, ncu --set full of the two stream kernels.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_r1c.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== bench c2 chunk"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/c2_err.log | tee gpurun_out/bench_c2_chunk.json | cut -c1-1700; tail -3 gpurun_out/c2_err.log
ls -la /tmp/mmbcache; free -g | head -2
echo "=== bench c2 window"; MMB_DETECT=window timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/c2_err.log | tee gpurun_out/bench_c2_window.json | cut -c1-400
echo "=== phases c2"; MMB_MIN_WARMUP=1 MMB_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2> gpurun_out/c2_phases.log | cut -c1-200
grep '^\[phase\]' gpurun_out/c2_phases.log | python -c "
import sys, collections
tot=collections.OrderedDict(); n=collections.Counter()
lines=[l for l in sys.stdin]
half=len(lines)//2
for l in lines[half:]:
    name=l[8:37].strip(); ms=float(l.split()[-2]); tot[name]=tot.get(name,0)+ms; n[name]+=1
for k,v in tot.items(): print('%-30s %9.2f ms x%d'%(k,v,n[k]))
"
grep '^\[sa\]' gpurun_out/c2_phases.log | tail -40
echo "=== ncu launch list c2 (one warm step)"
MMB_MIN_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 11300 --launch-count 11300 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/launches_c2.csv; tail -c 300 gpurun_out/ncu_list.log
for k in expand_members detect_chunk_kernel; do
  echo "=== ncu full $k"
  MMB_MIN_WARMUP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -f -o gpurun_out/prof_c2_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_$k.log | cut -c1-200
  ls -la gpurun_out/prof_c2_$k.ncu-rep
done
echo "=== done"
