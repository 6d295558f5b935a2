#!/bin/bash
# round 2, visit F (8 GPUs): partition-merge at N = 8 and 4 through torchrun (bench.py) and through the C++ host (mumemto_exec --gpus)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/f_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "cli_partition_merge" > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -n 3 gpurun_out/f_pytest.log
export MMB_CACHE_DIR=/tmp/mmb_cache
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/f_bench_n8.json 2> gpurun_out/f_bench_n8.err; echo "exit $?" >> gpurun_out/f_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_n4.json 2> gpurun_out/f_bench_n4.err; echo "exit $?" >> gpurun_out/f_bench_n4.err
# the drop-in executable on 8 GPUs: 8 partitions of the 64 x 16 Mbp collection written to local disk
python - <<'PY' > gpurun_out/f_cli_gpus8.log 2>&1
import os, subprocess, sys, time, hashlib
sys.path.insert(0, os.getcwd())
import bench
files = bench.SampleFiles(dict(bench.WORKLOADS["h64_16m"]))
exe = os.path.join("mumemto_b200", "bin", "mumemto_exec")
for flag in (["--gpus", "8"], ["--partitions", "8"]):
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-o", os.path.join(files.dir, "out" + flag[0][2:]), *flag, *files.paths], capture_output=True, env=dict(os.environ, MMB_CLI_TIMING="1"))
    dt = time.perf_counter() - t0
    print(flag, "rc", r.returncode, "wall", round(dt, 3), "s")
    print(r.stderr.decode(errors="replace")[-1500:])
a = open(os.path.join(files.dir, "outgpus.mums"), "rb").read(); b = open(os.path.join(files.dir, "outpartitions.mums"), "rb").read()
ta = open(os.path.join(files.dir, "outgpus.athresh"), "rb").read(); tb = open(os.path.join(files.dir, "outpartitions.athresh"), "rb").read()
print("8 GPUs vs 8 sequential partitions on one GPU: .mums equal", a == b, len(a), "bytes; .athresh equal", ta == tb, "md5", hashlib.md5(a).hexdigest())
PY
tail -n 4 gpurun_out/f_cli_gpus8.log
for f in n8 n4; do echo == $f; grep -a '^{' gpurun_out/f_bench_$f.json | head -c 400; tail -n 3 gpurun_out/f_bench_$f.err; done
