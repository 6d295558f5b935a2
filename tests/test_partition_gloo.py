"""Multi-rank host logic of the partition-merge path on CPU: world_size 2 and 3 over gloo, each rank running its partition through the
SIMT unit-test build, one all-gather, merge on every rank; the result must equal the oracle's `merge` and, as a set, the direct run on
the union of the genomes (reference README.md:132)."""
import os
import subprocess
import sys

import pytest

import util

ROOT = util.ROOT

WORKER = r'''
import os, sys, ctypes, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from mumemto_b200 import _binding, partition
from mumemto_b200._binding import Engine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
lib = _binding.declare(ctypes.CDLL(os.path.join({root!r}, "tests", "simt", "_build", "libmumemto_b200_simt.so")))
groups = {groups!r}
paths = [os.path.join({data!r}, "h%d.fa" % i) for i in groups[rank]]
e = Engine(_lib=lib)
e.load_fasta_files(paths)
e.find(merge=True, anchor_merge=True)
lengths, offsets, strands, thr = partition.partition_merge(e, dist, world, torch.device("cpu"))
text = e.merge_text()
buf = np.zeros(e.merge_text_size() + 7, dtype=np.uint8)
assert e.merge_text_into(buf) == len(text) and buf[:len(text)].tobytes() == text
open(os.path.join({out!r}, "merged_rank%d.mums" % rank), "wb").write(text)
thr.tofile(os.path.join({out!r}, "merged_rank%d.athresh" % rank))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("groups", [[[0, 1, 2], [0, 3]], [[0, 1], [0, 2], [0, 3]]])
def test_partition_merge_over_gloo(simt_lib, oracle_bin, tmp_path, groups):
    data = str(tmp_path / "data")
    paths = util.appendix_a(data, N=4, L=2500, seed=11, identical_ends=True, snp=0.02, width=70)
    world = len(groups)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, groups=groups, data=data, out=str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + world))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29510 + world), str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    names = []
    for i, idx in enumerate(groups):
        pre = str(tmp_path / f"o{i}")
        util.run_cli(oracle_bin, ["-o", pre, "-M", "-n", *[paths[j] for j in idx]])
        names.append(pre + ".mums")
    util.run_cli(oracle_bin, ["merge", *names, "-o", str(tmp_path / "om")])
    want_mums, want_thr = open(tmp_path / "om.mums", "rb").read(), open(tmp_path / "om.athresh", "rb").read()
    for rank in range(world):
        assert open(tmp_path / f"merged_rank{rank}.mums", "rb").read() == want_mums
        assert open(tmp_path / f"merged_rank{rank}.athresh", "rb").read() == want_thr
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "all"), "-M", "-n", *paths])
    assert sorted(want_mums.splitlines()) == sorted(open(tmp_path / "all.mums", "rb").read().splitlines())
    assert want_thr == open(tmp_path / "all.athresh", "rb").read()


SHARD_WORKER = r'''
import os, sys, ctypes, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from mumemto_b200 import _binding, partition
from mumemto_b200._binding import Engine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
lib = _binding.declare(ctypes.CDLL(os.path.join({root!r}, "tests", "simt", "_build", "libmumemto_b200_simt.so")))
paths = [os.path.join({data!r}, "h%d.fa" % i) for i in range({n})]
e = Engine(_lib=lib)
e.load_fasta_files(paths)
st = e.find(shard=(rank, world), **{kw!r})
assert st["route"] == 2 and st["pruned"]
out, n = partition.gather_slices(e, dist, world, torch.device("cpu"))
open(os.path.join({out!r}, "sharded_rank%d.out" % rank), "wb").write(out.numpy()[:n].tobytes())
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world,flags", [(2, []), (3, ["-f", "3", "-F", "100"])])
def test_sharded_collection_over_gloo(simt_lib, oracle_bin, tmp_path, world, flags):
    """SURVEY 8e(2) host logic at world_size 2 / 3: every rank holds the whole collection and finds the matches of its slice of the suffix
    order; one size exchange + one all-gather of the output slices; every rank ends up with the bytes of the single-process run."""
    data = str(tmp_path / "data")
    paths = util.appendix_a(data, N=5, L=3000, seed=13, identical_ends=True, snp=0.015, width=70)
    script = tmp_path / "worker.py"
    script.write_text(SHARD_WORKER.format(root=ROOT, data=data, out=str(tmp_path), n=len(paths), kw=util.flags_to_kwargs(flags)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29520 + world))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29530 + world), str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "all"), *flags, *paths])
    want = open(str(tmp_path / "all") + (".mems" if flags else ".mums"), "rb").read()
    assert len(want) > 100
    for rank in range(world):
        assert open(tmp_path / f"sharded_rank{rank}.out", "rb").read() == want
