#!/bin/bash
# round 2, visit R (2 GPUs): final multi-GPU check of the round-end build: two-GPU tests, the default N = 2 line, sharded C2 on two GPUs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gpus or shard or partition or device_input" > gpurun_out/r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r_pytest.log
tail -n 4 gpurun_out/r_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r_bench_n2.json 2> gpurun_out/r_bench_n2.err; echo "exit $?" >> gpurun_out/r_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --sharded --steps 3 --warmup 3 --no-cpu-baseline --kernel-table-rows 12 > gpurun_out/r_bench_c2_sharded_n2.json 2> gpurun_out/r_bench_c2_sharded_n2.err; echo "exit $?" >> gpurun_out/r_bench_c2_sharded_n2.err
for f in n2 c2_sharded_n2; do echo == $f; grep -a '^{' gpurun_out/r_bench_$f.json | head -c 300; tail -n 2 gpurun_out/r_bench_$f.err; done
