#!/bin/bash
# round 2, visit U (1 GPU): trigger scan ahead of the parse during the upload (end-to-end leg); the default bench line of the round-end build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/u_pytest.log
tail -n 3 gpurun_out/u_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python bench.py --steps 10 --warmup 3 --kernel-table-rows 60 --cli-e2e > gpurun_out/u_bench_c2.json 2> gpurun_out/u_bench_c2.err; echo "exit $?" >> gpurun_out/u_bench_c2.err
echo == c2; head -c 250 gpurun_out/u_bench_c2.json; tail -n 2 gpurun_out/u_bench_c2.err
