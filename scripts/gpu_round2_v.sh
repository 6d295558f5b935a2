#!/bin/bash
# round 2, visit V (1 GPU): 64 x 40 Mbp (text crosses 2^32 symbols) in MEM and partial mode against the reference's md5s of that collection
mkdir -p gpurun_out
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 600 python bench.py --workload h64_40m_c4 --steps 2 --warmup 3 --no-cpu-baseline --kernel-table-steps 0 > gpurun_out/v_bench_h64_40m_c4.json 2> gpurun_out/v_bench_h64_40m_c4.err; echo "exit $?" >> gpurun_out/v_bench_h64_40m_c4.err
cp /tmp/mmb_cache/h64_40m_c4_r0_w1.npy /tmp/mmb_cache/h64_40m_c3_r0_w1.npy; cp /tmp/mmb_cache/h64_40m_c4_r0_w1.npy.sizes.npy /tmp/mmb_cache/h64_40m_c3_r0_w1.npy.sizes.npy
timeout 600 python bench.py --workload h64_40m_c3 --steps 2 --warmup 3 --no-cpu-baseline --kernel-table-steps 0 > gpurun_out/v_bench_h64_40m_c3.json 2> gpurun_out/v_bench_h64_40m_c3.err; echo "exit $?" >> gpurun_out/v_bench_h64_40m_c3.err
for f in c4 c3; do echo == $f; head -c 230 gpurun_out/v_bench_h64_40m_$f.json; tail -n 2 gpurun_out/v_bench_h64_40m_$f.err; done
