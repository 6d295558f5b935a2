"""String-based merge: mumemto_b200.merge_mums against outputs of the reference's own mumemto/merge_mums.py.

CPU part: the merge step proper (`merge_tables`, `split_at_terminators`, file readers/writers) on the committed intermediate files.
GPU part: the whole workflow on our binaries (partition runs with -M on the device, bin/extract_mums, the MUMs-of-MUMs run on the device,
the merge) must write the reference's bytes.  Fixtures: tests/golden/string_merge/ (tests/golden/make_string_merge_golden.py)."""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import util
from mumemto_b200 import merge_mums as mm

G = os.path.join(util.GOLDEN, "string_merge")
sys.path.insert(0, util.GOLDEN)
CASES = ["two", "three", "reverse"]


def _table():
    with open(os.path.join(G, "table.json")) as f:
        return json.load(f)


def _read(path):
    with open(path, "rb") as f:
        return f.read()


@pytest.mark.parametrize("case", CASES)
def test_merge_step_matches_reference(case, tmp_path):
    """`-m mom.mums`: only the merge step runs (no engine, no extract_mums) -- .mums, .bumbl, .thresh, .thresh_rev byte-equal."""
    d = tmp_path / case
    shutil.copytree(os.path.join(G, case), d)
    k = len(_table()[case]["partitions"])
    for i in range(k):                                          # the merged .lengths is assembled from the partitions' files
        (d / f"p{i}.lengths").write_text(f"/x/h{i}.fa 10\n")
    for out, ref in (("out.mums", "merged.mums"), ("out_b.bumbl", "merged_m.bumbl")):
        a = mm.parse_arguments(["-m", str(d / "mom.mums"), *[str(d / f"p{i}.mums") for i in range(k)], "-o", str(d / out)])
        assert mm.main(a) == 0
        assert _read(d / out) == _read(d / ref)
        base = str(d / out).rsplit(".", 1)[0]
        assert _read(base + ".thresh") == _read(d / "merged.thresh")
        assert _read(base + ".thresh_rev") == _read(d / "merged.thresh_rev")
    assert (d / "out.lengths").read_text() == "\n".join(f"/x/h{i}.fa 10" for i in range(k))
    assert (d / "mom.mums").exists()                            # a file given with -m is the caller's: not cleaned up


def test_split_at_terminators_matches_reference_vectors():
    with open(os.path.join(G, "split_vectors.json")) as f:
        vectors = json.load(f)
    cut = 0
    for v in vectors:
        ends = np.cumsum(v["records"])
        bv = np.zeros(int(ends[-1]) + 1, bool)
        bv[ends - 1] = True
        l, s, d = mm.split_at_terminators(np.array([v["l"]], np.uint32), np.array([v["starts"]], np.int64), np.array([v["strands"]], bool), bv)
        got = [[int(a), [int(x) for x in b], [bool(x) for x in c]] for a, b, c in zip(l, s, d)]
        assert got == v["pieces"]
        cut += len(v["pieces"]) != 1 or v["pieces"][0][0] != v["l"]
    assert cut > 20                                             # the vectors do exercise the cutting


def test_table_round_trip_and_sort(tmp_path):
    rng = np.random.default_rng(0)
    n, k = 37, 5
    lengths = rng.integers(20, 5000, n).astype(np.uint32)
    starts = rng.integers(0, 1 << 40, (n, k)).astype(np.int64)
    strands = rng.integers(0, 2, (n, k)).astype(bool)
    for ext in (".mums", ".bumbl"):
        p = str(tmp_path / ("t" + ext))
        mm.write_table(p, lengths, starts, strands)
        l2, s2, d2 = mm.read_table(p)
        assert (l2 == lengths).all() and (s2 == starts).all() and (d2 == strands).all()
    l3, s3, _ = mm.sort_table((lengths, starts, strands))
    assert (np.diff(s3[:, 0]) >= 0).all() and sorted(l3.tolist()) == sorted(lengths.tolist())


def test_lengths_files(tmp_path):
    (tmp_path / "a.lengths").write_text("/d/h0.fa 100\n/d/h1.fa 90\n")
    (tmp_path / "b.lengths").write_text("/d/h0.fa * 100\n/d/h0.fa r1 60\n/d/h0.fa r2 40\n/d/h2.fa * 7\n/d/h2.fa r 7\n")
    for f in ("a", "b"):
        (tmp_path / (f + ".mums")).write_text("")
    a = mm.parse_arguments([str(tmp_path / "a.mums"), str(tmp_path / "b.mums"), "-o", str(tmp_path / "o")])
    mm.merge_lengths(a)                                         # mixed forms: single-record entries are rewritten
    assert (tmp_path / "o.lengths").read_text().split("\n")[:3] == ["/d/h0.fa * 100", "/d/h0.fa h0.fa 100", "/d/h1.fa * 90"]
    mm.merge_anchor_lengths(a)                                  # the anchor's entries of later partitions are dropped
    assert "/d/h0.fa r1 60" not in (tmp_path / "o.lengths").read_text()
    assert mm.record_lengths(str(tmp_path / "b.lengths")) == [[60, 40], [7]]
    with pytest.raises(ValueError):
        mm.record_lengths(str(tmp_path / "a.lengths"))


def test_extract_mums_tool(tmp_path):
    """bin/extract_mums (host tool, no device): the committed FASTA records of the reference's extract_mums, from .mums and .bumbl."""
    exe = os.path.join(util.ROOT, "mumemto_b200", "bin", "extract_mums")
    if not os.path.exists(exe):
        pytest.skip("bin/extract_mums not built")
    import make_string_merge_golden as gen
    paths, parts = gen.case_inputs(str(tmp_path), "two")
    for i, idx in enumerate(parts):
        shutil.copy(os.path.join(G, "two", f"p{i}.mums"), tmp_path / f"p{i}.mums")
        (tmp_path / f"p{i}.lengths").write_text("".join(f"{paths[j]} 8000\n" for j in idx))
        util.run_cli(exe, ["-m", tmp_path / f"p{i}.mums"])
        assert _read(tmp_path / f"p{i}_mums.fa") == _read(os.path.join(G, "two", f"p{i}_mums.fa"))
        mm.write_table(str(tmp_path / f"q{i}.bumbl"), *mm.read_table(str(tmp_path / f"p{i}.mums")))
        util.run_cli(exe, ["-m", tmp_path / f"q{i}.bumbl", "-l", tmp_path / f"p{i}.lengths", "-o", tmp_path / f"q{i}"])
        assert _read(tmp_path / f"q{i}.fa") == _read(os.path.join(G, "two", f"p{i}_mums.fa"))
    assert util.run_cli(exe, ["-m", tmp_path / "nothing.mums"], check=False).returncode == 1


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_string_merge_workflow_on_device(case, tmp_path):
    """Partitions with -M on the device -> python -m mumemto_b200.merge_mums (extract_mums, engine on the records, merge)."""
    import make_string_merge_golden as gen
    exe = os.path.join(util.ROOT, "mumemto_b200", "bin", "mumemto_exec")
    paths, parts = gen.case_inputs(str(tmp_path), case)
    golden = os.path.join(G, case)
    for i, idx in enumerate(parts):
        util.run_cli(exe, ["-o", tmp_path / f"p{i}", "-M", *[paths[j] for j in idx]])
        for ext in (".mums", ".thresh", ".thresh_rev"):
            assert _read(tmp_path / f"p{i}{ext}") == _read(os.path.join(golden, f"p{i}{ext}")), (i, ext)
    r = subprocess.run([sys.executable, "-m", "mumemto_b200.merge_mums", *[str(tmp_path / f"p{i}.mums") for i in range(len(parts))],
                        "-o", str(tmp_path / "merged")], cwd=util.ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    for ext in (".mums", ".thresh", ".thresh_rev"):
        assert _read(tmp_path / f"merged{ext}") == _read(os.path.join(golden, f"merged{ext}")), ext
    assert not (tmp_path / "p0_mums.fa").exists() and not (tmp_path / "merged_temp_merged.mums").exists()
