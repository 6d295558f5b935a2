#!/bin/bash
# round 2, visit D (2 GPUs): C++ multi-GPU host, torchrun partition-merge bench at N=2, one C5 partition with the mask threshold kernel, CLI timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/d_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "partition or merge or cli or pangenome or bench_parity" > gpurun_out/d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/d_pytest.log
tail -n 4 gpurun_out/d_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err; echo "exit $?" >> gpurun_out/d_bench_n2.err
timeout 600 python bench.py --workload c5 --merge-mode --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --kernel-table-rows 60 > gpurun_out/d_bench_c5p.json 2> gpurun_out/d_bench_c5p.err; echo "exit $?" >> gpurun_out/d_bench_c5p.err
timeout 900 python bench.py --steps 3 --warmup 3 --cli-e2e --no-cpu-baseline --kernel-table-rows 60 > gpurun_out/d_bench_c2.json 2> gpurun_out/d_bench_c2.err; echo "exit $?" >> gpurun_out/d_bench_c2.err
for f in n2 c5p c2; do echo == $f; head -c 300 gpurun_out/d_bench_$f.json; tail -n 3 gpurun_out/d_bench_$f.err; done
# ncu: launch list of one step at 64 x 16 Mbp, and --set full captures of the new per-rank kernels
export MMB_MIN_WARMUP=1
NCU_BENCH="python bench.py --workload h64_16m --route pfp --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --kernel-table-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_h64_16m.csv $NCU_BENCH > gpurun_out/d_ncu_list.log 2>&1
for k in classify_segments rank_words member_flags sort_dgroups_warp parse_lcp fasta_pass; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/d_ncu_$k $NCU_BENCH > gpurun_out/d_ncu_$k.log 2>&1
done
NCU_M="python bench.py --workload h64_16m --route pfp --merge-mode --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --kernel-table-steps 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:segment_thresholds -c 1 -o gpurun_out/d_ncu_segment_thresholds $NCU_M > gpurun_out/d_ncu_segthr.log 2>&1
ls -la gpurun_out/*.ncu-rep
