#!/bin/bash
# ncu --set full captures of the top kernels (one launch each) on the 64x16 Mbp workload, PFP route.
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_ncu.log) 2>&1
for k in expand_members detect_window_kernel; do
  echo "=== ncu full $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k python bench.py --workload h64_16m --route pfp --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_$k.log | cut -c1-200
  ls -la gpurun_out/prof_$k.ncu-rep
done
echo "=== done"
