"""The CPU oracle (oracle/oracle.cpp) is pinned against (1) the committed golden vectors, which were produced by the reference's own
sources compiled unmodified (tests/golden/make_golden.py), and (2) that reference build run live when
it is available.  CPU only."""
import os

import pytest

import util

T2 = dict(N=5, L=20000, seed=42, identical_ends=True)
T2_FLAGS = {"strict": [], "partial": ["-k", "-1"], "mem": ["-f", "3", "-F", "100"], "bumbl": ["-b"], "partial_bumbl": ["-k", "-1", "-b"],
            "anchor": ["-M", "-n"], "string_merge": ["-M"], "arrays": ["-A"], "norc": ["-r"], "short": ["-l", "12"],
            "mem2": ["-k", "3", "-f", "2"], "memF": ["-f", "0", "-F", "7", "-k", "2"]}


@pytest.fixture(scope="module")
def t2(tmp_path_factory):
    d = tmp_path_factory.mktemp("t2")
    paths = util.appendix_a(str(d), **T2)
    table = util.load_golden_table()["t2"]
    for p in paths:
        assert util.md5_file(p) == table["inputs"][os.path.basename(p)], "input generator drifted from the golden inputs"
    return str(d), paths, table


@pytest.mark.parametrize("name", sorted(T2_FLAGS))
def test_oracle_matches_golden_t2(oracle_bin, t2, name):
    d, paths, table = t2
    util.run_cli(oracle_bin, ["-o", os.path.join(d, "o_" + name), *T2_FLAGS[name], *paths])
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "o_" + name)).items()}
    assert got == table["outputs"][name]["md5"]


def test_golden_files_match_table(t2):
    _, _, table = t2
    for name in ("strict", "partial", "mem", "short", "merge2", "merge3"):
        ext = ".mems" if name.startswith("mem") and not name.startswith("merge") else ".mums"
        assert util.md5_file(os.path.join(util.GOLDEN, "t2", name + ext)) == table["outputs"][name]["md5"][ext]


@pytest.mark.parametrize("tag", ["merge2", "merge3"])
def test_oracle_merge_matches_golden(oracle_bin, t2, tag):
    d, paths, table = t2
    names = []
    for i, idx in enumerate(table["outputs"][tag]["partitions"]):
        pre = os.path.join(d, f"om_{tag}_{i}")
        util.run_cli(oracle_bin, ["-o", pre, "-M", "-n", *[paths[j] for j in idx]])
        names.append(pre + ".mums")
    util.run_cli(oracle_bin, ["merge", *names, "-o", os.path.join(d, "om_" + tag)])
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "om_" + tag)).items()}
    assert got == table["outputs"][tag]["md5"]


def test_oracle_toy_mem_quirk(oracle_bin, tmp_path):
    """mumemto_library/README.md:53,58 inputs; shows the last-occurrence +1 quirk of write_mem (include/mem_finder.hpp:229 vs :248)."""
    a = util.write_fasta(str(tmp_path / "a.fa"), [("a", "ACGTACGT")])
    b = util.write_fasta(str(tmp_path / "b.fa"), [("b", "TACGTAAA")])
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "toy"), "-l", "3", "-f", "2", a, b])
    toy = util.load_golden_table()["toy"]
    assert open(tmp_path / "toy.mems").read() == toy["mems"]
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "toy2"), "-l", "3", a, b])
    assert open(tmp_path / "toy2.mums").read() == ""      # doc 0 is its own reverse complement: no MUM survives


def test_oracle_matches_golden_c1_strict(oracle_bin, tmp_path):
    """C1 = BASELINE configs[0]: 5 x 1 Mbp, 1 % SNPs."""
    paths = util.appendix_a(str(tmp_path), N=5, L=1000000, seed=1)
    table = util.load_golden_table()["t7"]
    assert util.md5_file(paths[0]) == table["inputs"]["h0.fa"]
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "o"), *paths])
    assert util.md5_file(str(tmp_path / "o.mums")) == table["outputs"]["strict"]["md5"][".mums"]


def test_reference_build_matches_golden(ref_dir, t2):
    """The live reference build (when /root/reference or a prebuilt oracle/_ref is present) still reproduces the committed vectors."""
    d, paths, table = t2
    for name in ("strict", "mem", "anchor"):
        util.run_cli(os.path.join(ref_dir, "mumemto_exec"), ["-o", os.path.join(d, "r_" + name), *T2_FLAGS[name], *paths])
        got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "r_" + name)).items()}
        assert got == table["outputs"][name]["md5"]


def _edge_inputs(d):
    import numpy as np
    rng = np.random.default_rng(5)
    core = util.random_dna(rng, 900)
    docs = {
        "multi": [("c1", core[:300]), ("c2", core[300:620]), ("c3", core[620:])],                 # multi-record file (Q9)
        "lower": [("x", util.mutate(rng, core, 0.02).lower())],                                      # lower case input
        "iupac": [("y", util.mutate(rng, core, 0.01)[:400] + "NNNNNRYKMSWBDHVN" + core[416:])],      # N run + IUPAC codes
        "plain": [("z", util.mutate(rng, core, 0.03))],
    }
    paths = []
    for name, recs in docs.items():
        paths.append(util.write_fasta(os.path.join(d, name + ".fa"), recs, width=50, newline="\r\n" if name == "iupac" else "\n"))
    return paths


@pytest.mark.parametrize("flags", [[], ["-k", "-1"], ["-f", "2"], ["-r"], ["-M", "-n"], ["-l", "8", "-f", "4", "-F", "9"]])
def test_oracle_equals_reference_on_edge_inputs(oracle_bin, ref_dir, tmp_path, flags):
    """Multi-record FASTA, CRLF, lower case, N and IUPAC symbols: oracle restatement == reference build, byte for byte."""
    paths = _edge_inputs(str(tmp_path))
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "o"), *flags, *paths])
    util.run_cli(os.path.join(ref_dir, "mumemto_exec"), ["-o", str(tmp_path / "r"), *flags, *paths])
    assert util.outputs(str(tmp_path / "o")) == util.outputs(str(tmp_path / "r"))
    assert open(tmp_path / "o.lengths").read() == open(tmp_path / "r.lengths").read()
