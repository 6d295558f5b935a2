#!/bin/bash
# round 2, visit W (2 GPUs): the default N = 2 line of the round-end build (partition-merge, end-to-end leg with the trigger scan ahead)
mkdir -p gpurun_out
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --kernel-table-steps 0 > gpurun_out/w_bench_n2.json 2> gpurun_out/w_bench_n2.err; echo "exit $?" >> gpurun_out/w_bench_n2.err
grep -a '^{' gpurun_out/w_bench_n2.json | head -c 300; tail -n 2 gpurun_out/w_bench_n2.err
