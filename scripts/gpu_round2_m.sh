#!/bin/bash
# round 2, visit M (2 GPUs): GPU tests incl. the sharded-collection mode on two devices; the default N = 2 bench line
# (partition-merge) and the strong-scaling line of ONE C2 collection over two GPUs (--sharded)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/m_gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_pytest.log
tail -n 6 gpurun_out/m_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/m_bench_n2.json 2> gpurun_out/m_bench_n2.err; echo "exit $?" >> gpurun_out/m_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --sharded --steps 5 --warmup 3 --no-cpu-baseline --kernel-table-rows 30 > gpurun_out/m_bench_c2_sharded_n2.json 2> gpurun_out/m_bench_c2_sharded_n2.err; echo "exit $?" >> gpurun_out/m_bench_c2_sharded_n2.err
for f in n2 c2_sharded_n2; do echo == $f; grep -a '^{' gpurun_out/m_bench_$f.json | head -c 400; tail -n 3 gpurun_out/m_bench_$f.err; done
