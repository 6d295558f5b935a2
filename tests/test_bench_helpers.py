"""bench.py's measurement helpers on CPU: workload generators (C1 must be the golden collection), the sampled output verifier, and the
legitimacy of the reference arm's partition mode (the merged result must be the direct result as a set)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import bench
import util

REF = os.path.join(util.ROOT, "oracle", "_ref", "mumemto_exec")
REF_MERGE = os.path.join(util.ROOT, "oracle", "_ref", "anchor_merge")


def test_c1_workload_is_the_golden_collection():
    """bench.py --workload c1 = SURVEY.md Appendix A t7: md5 of h0.fa / h4.fa as recorded there."""
    imgs = bench.make_collection(bench.WORKLOADS["c1"], seed=1)
    assert len(imgs) == 5
    assert hashlib.md5(imgs[0].tobytes()).hexdigest() == "35a1e79a633c9e3f0f907447347d9cc6"
    assert hashlib.md5(imgs[4].tobytes()).hexdigest().startswith("b3789bea")


def test_fasta_image_layout():
    codes = np.arange(205, dtype=np.uint8) % 4
    im = bench.fasta_image("x", codes, width=80).tobytes()
    lines = im.split(b"\n")
    assert lines[0] == b">x" and [len(l) for l in lines[1:]] == [80, 80, 45, 0]
    assert b"".join(lines[1:]) == bytes(b"ACGT"[c] for c in codes)


def test_pangenome_generator_is_deterministic_and_shares_variants():
    spec = dict(N=6, L=30000, kind="pangenome")
    a, b = bench.make_collection(spec, seed=5), bench.make_collection(spec, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # partitions generated for different ranks share document 0 (the anchor) and nothing else
    p0 = bench.make_collection(spec, seed=1, anchor_seed=12345)
    p1 = bench.make_collection(spec, seed=1001, anchor_seed=12345)
    assert np.array_equal(p0[0], p1[0]) and not np.array_equal(p0[1], p1[1])


def test_verify_sample_accepts_the_oracle_and_rejects_a_wrong_line(oracle_bin, tmp_path):
    spec = dict(N=6, L=40000, kind="pangenome")
    imgs = bench.make_collection(spec, seed=3)
    paths = []
    for i, im in enumerate(imgs):
        p = str(tmp_path / f"h{i}.fa"); im.tofile(p); paths.append(p)
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "o"), *paths])
    text = open(tmp_path / "o.mums", "rb").read()
    assert text.count(b"\n") > 50
    assert bench.verify_sample(imgs, text, how_many=200) > 50
    lines = text.split(b"\n")
    f = lines[0].split(b"\t")
    offs = f[1].split(b",")
    offs[1] = str(int(offs[1]) + 3).encode()                 # shift one occurrence of the first line
    lines[0] = b"\t".join([f[0], b",".join(offs), f[2]])
    with pytest.raises(AssertionError):
        bench.verify_sample(imgs, b"\n".join(lines), how_many=5)


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(REF_MERGE)), reason="oracle/_ref is not built")
def test_reference_partition_mode_gives_the_direct_result(tmp_path):
    """The reference arm's many-core mode (partitions sharing document 0 as concurrent processes + anchor_merge) is the reference's own
    workflow and must produce the multi-MUMs of the direct run, as a set (reference README.md:126-134)."""
    spec = dict(N=9, L=30000, kind="pangenome")
    imgs = bench.make_collection(spec, seed=1)
    paths = []
    for i, im in enumerate(imgs):
        p = str(tmp_path / f"h{i}.fa"); im.tofile(p); paths.append(p)
    procs = 3
    groups = [[paths[0]] + paths[1 + g::procs] for g in range(procs)]
    for g, grp in enumerate(groups):
        subprocess.run([REF, "-o", str(tmp_path / f"part{g}"), "-M", "-n", *grp], check=True, capture_output=True)
    subprocess.run([REF_MERGE, *[str(tmp_path / f"part{g}.mums") for g in range(procs)], "-o", str(tmp_path / "merged")], check=True, capture_output=True)
    subprocess.run([REF, "-o", str(tmp_path / "direct"), *paths], check=True, capture_output=True)
    order = [0] + [i for g in range(procs) for i in range(1 + g, len(paths), procs)]       # column order of the merged table

    def rows(path, cols):
        out = set()
        for line in open(path):
            l, o, s = line.rstrip("\n").split("\t")
            o, s = o.split(","), s.split(",")
            out.add((l, tuple(o[c] for c in cols), tuple(s[c] for c in cols)))
        return out

    inv = [order.index(i) for i in range(len(paths))]
    assert rows(tmp_path / "merged.mums", inv) == rows(tmp_path / "direct.mums", range(len(paths)))
    files = bench.SampleFiles(dict(spec, L=8000))
    secs, groups, mums, athresh = bench.run_reference_partitioned(REF, REF_MERGE, files, 3)
    files.close()
    assert secs > 0 and len(groups) == 3 and len(athresh) == 2 * 8001 and mums.count(b"\n") > 10


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(REF_MERGE)), reason="oracle/_ref is not built")
@pytest.mark.parametrize("workload_flags", ["strict", "partial", "mem"])
def test_parity_sample_check_runs_and_catches_a_difference(simt_lib, workload_flags, monkeypatch):
    """bench.py's off-the-clock parity check (reference output vs engine output on the same sample, byte for byte; partitions ->
    anchor_merge vs the device-input merge) driven through the CPU unit-test build of the kernels."""
    import torch
    from mumemto_b200 import Engine
    spec = dict(N=5, L=30000, kind="pangenome", flagset=workload_flags, desc="t")
    eng = Engine(0, _lib=simt_lib)
    base, parity = bench.cpu_baseline_and_parity(eng, torch.device("cpu"), spec, sample_L=3000, procs=2)
    assert parity["result"] == "ok" and parity["checks"]["direct"]["lines"] > 5
    assert ("partitions_anchor_merge" in parity["checks"]) == (workload_flags == "strict")
    assert base["value"] > 0 and base["extrapolated"]
    monkeypatch.setattr(bench, "run_reference", lambda *a, **k: (1.0, b"20\t1,2\t+,+\n"))
    with pytest.raises(AssertionError):
        bench.cpu_baseline_and_parity(eng, torch.device("cpu"), spec, sample_L=3000, procs=2)
    eng.close()
