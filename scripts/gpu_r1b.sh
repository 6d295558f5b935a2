#!/bin/bash
# Round-1 re-entry visit: parity tests, C2 bench line, per-phase host timings, ncu launch list of a C2 step, ncu --set full of the top kernels.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_r1b.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; free -g | head -2
echo "=== build+smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" || echo SMOKE_FAILED
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -20
echo "=== bench c2 (default)"; timeout 1800 python bench.py 2> gpurun_out/c2_err.log | tee gpurun_out/bench_c2.json | cut -c1-2500; tail -5 gpurun_out/c2_err.log
echo "=== phases c2"; MMB_DEBUG=1 timeout 900 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/c2_phases.log | cut -c1-300
grep phase gpurun_out/c2_phases.log | tail -120 | awk '{a[$2" "$3" "$4" "$5]+=$(NF-1); n[$2" "$3" "$4" "$5]++} END {for (k in a) printf "%-40s %10.2f ms  x%d\n", k, a[k], n[k]}' | sort -k1,1
echo "=== ncu launch list c2"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/launches_c2.csv; tail -c 400 gpurun_out/ncu_list.log
for k in expand_members detect_window_kernel sort_groups_warp; do
  echo "=== ncu full $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -f -o gpurun_out/prof_c2_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_$k.log | cut -c1-200
  ls -la gpurun_out/prof_c2_$k.ncu-rep
done
echo "=== done"
