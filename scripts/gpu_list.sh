#!/bin/bash
# ncu launch list (gpu__time_duration only) of a short C2 run: 1 warm-up + 1 timed + 1 end-to-end step.
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_list.log) 2>&1
export MMB_CACHE_DIR=/tmp/mmbcache
W=${1:-c2}
MMB_MIN_WARMUP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/launches_$W.csv; tail -c 300 gpurun_out/ncu_list.log
echo "=== done"
