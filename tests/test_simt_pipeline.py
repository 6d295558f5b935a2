"""Kernel logic of the engine, executed through the SIMT unit-test build (tests/simt) and compared byte-for-byte with the CPU oracle.
CPU only: this checks index arithmetic, halo handling and host orchestration of the sm_100a sources before they reach a GPU; the
parity tests proper are tests/test_gpu_parity.py (-m gpu), which run the nvcc build through the same C ABI."""
import os

import numpy as np
import pytest

import util
from mumemto_b200 import api

FLAGSETS = {"strict": [], "partial": ["-k", "-1"], "mem": ["-f", "3", "-F", "100"], "bumbl": ["-b"], "partial_bumbl": ["-k", "-2", "-b"],
            "anchor": ["-M", "-n"], "arrays": ["-A"], "norc": ["-r"], "short": ["-l", "9"], "mem2": ["-k", "3", "-f", "2"],
            "memF": ["-f", "0", "-F", "7", "-k", "2"], "k1": ["-k", "1"], "kbig": ["-k", "99"], "kneg": ["-k", "-99"], "F1": ["-F", "1", "-f", "2"],
            "Fneg": ["-f", "2", "-F", "-1", "-k", "2"], "f5": ["-f", "5", "-l", "10"]}


def both(simt_lib, oracle_bin, tmp, paths, flags):
    o, g = os.path.join(tmp, "o"), os.path.join(tmp, "g")
    util.run_cli(oracle_bin, ["-o", o, *flags, *paths])
    api.find_files(paths, g, _lib=simt_lib, **util.flags_to_kwargs(flags))
    assert util.outputs(g) == util.outputs(o)
    assert open(g + ".lengths").read() == open(o + ".lengths").read()
    return util.outputs(g)


@pytest.fixture(scope="module")
def small(tmp_path_factory):
    d = tmp_path_factory.mktemp("small")
    return util.appendix_a(str(d), N=4, L=2500, seed=11, identical_ends=True, snp=0.02, width=70)


@pytest.mark.parametrize("name", sorted(FLAGSETS))
def test_simt_equals_oracle(simt_lib, oracle_bin, small, tmp_path, name):
    outs = both(simt_lib, oracle_bin, str(tmp_path), small, FLAGSETS[name])
    assert any(len(b) > 0 for b in outs.values())


def test_simt_edge_inputs(simt_lib, oracle_bin, tmp_path):
    """Multi-record FASTA, CRLF line ends, lower case, N runs, IUPAC codes, blank lines, leading junk before the first header."""
    rng = np.random.default_rng(5)
    core = util.random_dna(rng, 700)
    p = []
    p.append(util.write_fasta(str(tmp_path / "multi.fa"), [("c1 desc", core[:250]), ("c2", core[250:480]), ("c3", core[480:])], width=50))
    p.append(util.write_fasta(str(tmp_path / "lower.fasta"), [("x", util.mutate(rng, core, 0.02).lower())], width=61))
    p.append(util.write_fasta(str(tmp_path / "iupac.fna"), [("y", util.mutate(rng, core, 0.01)[:300] + "NNNNNRYKMSWBDHVNU" + core[317:])], width=50, newline="\r\n"))
    with open(tmp_path / "odd.fa", "w") as f:
        f.write("junk before header\n\n>z some comment > with marker\n" + core[:100] + "\n\n" + core[100:333] + "\n>empty\n>tail\n" + core[333:])   # no final newline
    p.append(str(tmp_path / "odd.fa"))
    for flags in ([], ["-f", "2"], ["-k", "-1", "-l", "10"], ["-r"]):
        both(simt_lib, oracle_bin, str(tmp_path), p, flags)


def _long_line_inputs(tmp_path, rng, n):
    """FASTA shapes where whole 4 KiB chunks hold no newline: single-line sequences, a header longer than a chunk, CRLF on long lines,
    records that end exactly at / one byte around a chunk boundary, a chunk that is all header followed by a marker-free chunk."""
    core = util.random_dna(rng, n)
    p = []
    p.append(util.write_fasta(str(tmp_path / "oneline.fa"), [("a", core)], width=10 ** 9))
    p.append(util.write_fasta(str(tmp_path / "crlf_long.fa"), [("b", util.mutate(rng, core, 0.01)[: n // 2]), ("b2", core[n // 2:])], width=9000, newline="\r\n"))
    with open(tmp_path / "longhdr.fa", "w") as f:          # header of 5000 characters with markers inside, then lines of 4095 / 4096 / 4097 bytes
        m = util.mutate(rng, core, 0.02)
        f.write(">" + "h>@+ " * 1000 + "\n" + m[:4095] + "\n" + m[4095:8191] + "\n" + m[8191:12288] + "\n>x\n" + m[12288:])
    p.append(str(tmp_path / "longhdr.fa"))
    with open(tmp_path / "edge.fa", "w") as f:             # first record ends so that the next header starts at byte 4096 exactly
        m = util.mutate(rng, core, 0.01)
        head = ">e\n"
        f.write(head + m[:4096 - len(head) - 1] + "\n" + ">e2 " + "d" * 4090 + "\n" + m[4096:] + "\n\n")
    p.append(str(tmp_path / "edge.fa"))
    return p


def test_simt_text_long_lines_and_headers(simt_lib, oracle_bin, tmp_path):
    p = _long_line_inputs(tmp_path, np.random.default_rng(8), 14000)
    for flags in ([], ["-r", "-f", "2"]):
        both(simt_lib, oracle_bin, str(tmp_path), p, flags)


def test_simt_gz_and_duplicates(simt_lib, oracle_bin, small, tmp_path):
    import gzip, shutil
    gz = str(tmp_path / "h0.fa.gz")
    with open(small[0], "rb") as a, gzip.open(gz, "wb") as b:
        shutil.copyfileobj(a, b)
    both(simt_lib, oracle_bin, str(tmp_path), [gz, small[1], small[1], small[2]], [])       # duplicate path is dropped


def test_simt_last_rank_interval_is_lost(simt_lib, oracle_bin, tmp_path):
    """Q6: no flush at end of stream -- a MUM whose interval reaches the last suffix-array rank is never reported."""
    rng = np.random.default_rng(3)
    S = "TTTTGACGTCAGCTAGCTAGCATCGA"
    a = util.random_dna(rng, 50, "ACG") + "G" + S + "A" + util.random_dna(rng, 50, "ACG")
    b = util.random_dna(rng, 50, "ACG") + "A" + S + "C" + util.random_dna(rng, 50, "ACG")
    p = [util.write_fasta(str(tmp_path / "a.fa"), [("a", a)]), util.write_fasta(str(tmp_path / "b.fa"), [("b", b)])]
    outs = both(simt_lib, oracle_bin, str(tmp_path), p, ["-r"])
    assert outs[".mums"] == b""


def test_simt_threshold_saturation(simt_lib, oracle_bin, tmp_path):
    """-M with LCP values above 65535: thresholds saturate at 65535 (include/mem_finder.hpp:299,328).  X occurs once per document and
    its 66000-symbol prefix once more in document 0, so the MUM intervals of X's suffixes have a neighbour sharing up to 66000."""
    rng = np.random.default_rng(9)
    X = util.random_dna(rng, 70000)
    docs = [X + "A" + X[:66000], X]
    p = [util.write_fasta(str(tmp_path / f"d{i}.fa"), [(f"d{i}", s)], width=80) for i, s in enumerate(docs)]
    outs = both(simt_lib, oracle_bin, str(tmp_path), p, ["-M", "-n", "-r"])
    t = np.frombuffer(outs[".athresh"], dtype=np.uint16)
    assert t.max() == 65535 and (t == 65535).sum() > 100


def test_simt_library_mirror(simt_lib, oracle_bin, tmp_path):
    """mumemto.mum / mumemto.mem of the reference's Python bindings, on the README toy and on random documents."""
    r = api.mem([["ACGTACGT"], ["TACGTAAA"]], min_match_len=3, max_doc_freq=2, _lib=simt_lib)
    assert len(r) == 2
    l, off, ids, st = r.match_at(0)
    assert (l, off.tolist(), ids.tolist(), st.tolist()) == (5, [0, 1, 0, 4], [1, 1, 0, 0], [0, 1, 1, 0])      # -,+,+,- with the +1 quirk
    assert len(api.mum([["ACGTACGT"], ["TACGTAAA"]], min_match_len=3, _lib=simt_lib)) == 0
    with pytest.raises(ValueError):
        api.mem([["ACGT"], ["ACGT"]], max_doc_freq=1, _lib=simt_lib)
    # multi-record documents against the oracle's text output
    rng = np.random.default_rng(2)
    core = util.random_dna(rng, 1500)
    docs = [[core[:700], core[700:]], [util.mutate(rng, core, 0.02)], [util.mutate(rng, core, 0.02)[:800], util.mutate(rng, core, 0.01)[800:]]]
    paths = [util.write_fasta(str(tmp_path / f"d{i}.fa"), [(f"r{j}", s) for j, s in enumerate(d)]) for i, d in enumerate(docs)]
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "o"), "-k", "2", *paths])
    res = api.mum(docs, num_distinct=2, _lib=simt_lib)
    lines = []
    for i in range(len(res)):
        l, off, st = res.match_at(i)
        lines.append(f"{l}\t" + ",".join("" if o < 0 else str(o) for o in off) + "\t" + ",".join("" if o < 0 else ("+" if s else "-") for o, s in zip(off, st)) + "\n")
    assert "".join(lines) == open(tmp_path / "o.mums").read()
    assert res.record_lengths == [[700, 800], [1500], [800, 700]]


def test_simt_errors(simt_lib, tmp_path):
    from mumemto_b200 import MumemtoError
    p = util.appendix_a(str(tmp_path), N=2, L=300, seed=1)
    with pytest.raises(MumemtoError, match="Merging not available for partial"):
        api.find_files([p[0], p[1], util.write_fasta(str(tmp_path / "c.fa"), [("c", "ACGT" * 50)])], str(tmp_path / "x"), num_distinct_docs=-1, merge=True, _lib=simt_lib)
    with pytest.raises(MumemtoError, match="Merging not available for multi-MEMs"):
        api.find_files(p, str(tmp_path / "x"), rare_freq=2, merge=True, _lib=simt_lib)
    with pytest.raises(MumemtoError, match="Multiple FASTA inputs required"):
        api.find_files([p[0], p[0]], str(tmp_path / "x"), _lib=simt_lib)
    open(tmp_path / "empty.fa", "w").write(">only a header\n")
    with pytest.raises(MumemtoError, match="Empty input file"):
        api.find_files([p[0], str(tmp_path / "empty.fa")], str(tmp_path / "x"), _lib=simt_lib)
    with pytest.raises(MumemtoError, match="one view"):                  # no frequency limit at all: needs the whole (pruned) stream in one view
        api.find_files(p, str(tmp_path / "x"), rare_freq=0, max_mem_freq=0, route="pfp", no_prune=True, _lib=simt_lib)


@pytest.mark.parametrize("parts", [[[0, 1, 2], [0, 3]], [[0, 1], [0, 2], [0, 3]], [[0, 3, 1], [0, 2]]])
def test_simt_anchor_merge(simt_lib, oracle_bin, small, tmp_path, parts):
    simt_lib.simt_set_block_order(len(parts) % 3)
    """Partition -> merge (row J): the device fold equals the oracle's fold, and (as a set) the direct run on the union."""
    names_o, names_g = [], []
    for i, idx in enumerate(parts):
        sel = [small[j] for j in idx]
        o, g = str(tmp_path / f"o{i}"), str(tmp_path / f"g{i}")
        util.run_cli(oracle_bin, ["-o", o, "-M", "-n", *sel])
        api.find_files(sel, g, merge=True, anchor_merge=True, _lib=simt_lib)
        assert util.outputs(g) == util.outputs(o)
        names_o.append(o + ".mums"); names_g.append(g + ".mums")
    util.run_cli(oracle_bin, ["merge", *names_o, "-o", str(tmp_path / "om")])
    api.merge_files(names_g, str(tmp_path / "gm"), _lib=simt_lib)
    assert util.outputs(str(tmp_path / "gm")) == util.outputs(str(tmp_path / "om"))
    if sorted(j for idx in parts for j in idx[1:]) == [1, 2, 3] and [j for idx in parts for j in idx[1:]] == [1, 2, 3]:
        util.run_cli(oracle_bin, ["-o", str(tmp_path / "all"), "-M", "-n", *small])
        assert sorted(open(tmp_path / "gm.mums").read().splitlines()) == sorted(open(tmp_path / "all.mums").read().splitlines())
        assert open(tmp_path / "gm.athresh", "rb").read() == open(tmp_path / "all.athresh", "rb").read()


@pytest.mark.parametrize("minlen", [9, 31])
def test_simt_anchor_merge_ignores_min_match_len(simt_lib, oracle_bin, small, tmp_path, minlen):
    """Quirk Q7: the merge keeps a match iff it is longer than the next-best threshold and at least 20 long -- hard-coded
    (src/merge_candidates.cpp:141), whatever -l the partitions were computed with.  Checked against the reference's own anchor_merge
    when oracle/_ref is built, else the oracle's fold."""
    ref_merge = os.path.join(util.ROOT, "oracle", "_ref", "anchor_merge")
    names = []
    for i, idx in enumerate(([0, 1, 2], [0, 3])):
        g = str(tmp_path / f"g{i}")
        api.find_files([small[j] for j in idx], g, merge=True, anchor_merge=True, min_match_len=minlen, _lib=simt_lib)
        names.append(g + ".mums")
    api.merge_files(names, str(tmp_path / "gm"), _lib=simt_lib)
    if os.path.exists(ref_merge):
        import subprocess
        subprocess.run([ref_merge, *names, "-o", str(tmp_path / "om")], check=True, capture_output=True)
    else:
        util.run_cli(oracle_bin, ["merge", *names, "-o", str(tmp_path / "om")])
    got = util.outputs(str(tmp_path / "gm"))
    assert got == util.outputs(str(tmp_path / "om"))
    lens = [int(l.split(b"\t")[0]) for l in got[".mums"].splitlines()]
    assert lens and min(lens) >= 20
    if minlen < 20:         # the partitions did hold shorter matches
        assert any(int(l.split("\t")[0]) < 20 for n in names for l in open(n))


@pytest.mark.parametrize("block_order", [0, 2])
def test_simt_repeats_all_group_size_classes(simt_lib, oracle_bin, tmp_path, block_order):
    """Low-complexity and tandem-repeat inputs: the doubling rounds see groups of every size class (sub-warp ... warp x8, block sort,
    device-wide fallback above 4096), and matches with many occurrences exercise -f / -F limits."""
    rng = np.random.default_rng(17)
    unit = util.random_dna(rng, 37)
    body = util.random_dna(rng, 900)
    docs = [
        "A" * 700 + body + unit * 50 + "ACGT" * 40 + body[:400],
        "A" * 300 + util.mutate(rng, body, 0.01) + unit * 30 + "C" * 150 + body[200:600],
        util.mutate(rng, body, 0.02) + "AC" * 400 + unit * 20 + "A" * 100,
    ]
    p = [util.write_fasta(str(tmp_path / f"d{i}.fa"), [(f"d{i}", s)], width=77) for i, s in enumerate(docs)]
    simt_lib.simt_set_block_order(block_order)       # real grids run their blocks in no particular order
    try:
        for flags in ([], ["-f", "4", "-k", "2", "-l", "15"], ["-r", "-f", "2"]) if block_order else ([], ["-k", "2"], ["-f", "4", "-k", "2", "-l", "15"], ["-f", "0", "-F", "40", "-k", "2"], ["-r", "-f", "2"]):
            both(simt_lib, oracle_bin, str(tmp_path), p, flags)
    finally:
        simt_lib.simt_set_block_order(0)


@pytest.mark.parametrize("flags", [[], ["-f", "3", "-k", "2"], ["-M", "-n"]])
def test_simt_walk_kernel_fallback(simt_lib, oracle_bin, small, tmp_path, flags, monkeypatch):
    """The leftward-walk detector (used when more than 64 interval sizes can qualify) must agree with the windowed one."""
    monkeypatch.setenv("MMB_FORCE_WALK", "1")
    both(simt_lib, oracle_bin, str(tmp_path), small, flags)


def test_simt_unbounded_frequency(simt_lib, oracle_bin, small, tmp_path):
    """-f 0 without -F: LCP intervals of any size; includes a collection with a 1200-fold repeat."""
    rng = np.random.default_rng(17)
    unit, body = util.random_dna(rng, 37), util.random_dna(rng, 3000)
    docs = ["A" * 900 + body + unit * 40 + "ACGT" * 100 + body[:700] + "G" * 450,
            "A" * 300 + util.mutate(rng, body, 0.01) + unit * 26 + "C" * 150 + body[200:900],
            util.mutate(rng, body, 0.02) + "AC" * 270 + unit * 12 + "A" * 100]
    rep = [util.write_fasta(str(tmp_path / f"r{i}.fa"), [(f"r{i}", s)], width=77) for i, s in enumerate(docs)]
    for name, paths in (("small", small), ("rep", rep)):
        for i, flags in enumerate((["-f", "0"], ["-f", "0", "-k", "2"], ["-f", "0", "-k", "2", "-l", "12", "-r"])):
            os.makedirs(str(tmp_path / f"{name}{i}"), exist_ok=True)
            out = both(simt_lib, oracle_bin, str(tmp_path / f"{name}{i}"), paths, flags)
            if name == "rep" and i == 0:
                assert max(len(l.split(b"\t")[1].split(b",")) for l in out[".mems"].splitlines()) > 1000
            if i < 1:       # the same on the prefix-free-parse route: the pruned stream in one view (section 8(f).4 on the tiled route)
                st = both_pfp(simt_lib, oracle_bin, str(tmp_path / f"{name}{i}"), paths, flags, tile_ranks=0)
                assert st["pruned"] and st["pfp_tiles"] == 1


def test_simt_trim_option(simt_lib, oracle_bin, small, tmp_path):
    """--trim (off by default): the reference's `mumemto trim` post-processing (mumemto/trim_mums.py:51-86) applied on the device: rows that
    run past the end of a sequence are shortened by their largest overhang, rows shorter than the minimum afterwards are dropped.  Checked
    against a numpy restatement on the untrimmed output; t2-like documents with identical ends hold such a MUM."""
    paths = util.appendix_a(str(tmp_path / "t"), N=4, L=3000, seed=42, identical_ends=True)
    for route in ("direct", "pfp"):
        a, b = str(tmp_path / f"plain_{route}"), str(tmp_path / f"trim_{route}")
        api.find_files(paths, a, route=route, _lib=simt_lib)
        api.find_files(paths, b, route=route, trim_min_len=20, _lib=simt_lib)
        want, changed = [], 0
        for line in open(a + ".mums"):
            l, o, sg = line.rstrip("\n").split("\t")
            over = max(int(x) + int(l) - 3000 for x in o.split(",") if x != "")
            nl = int(l) - over if over > 0 else int(l)
            changed += over > 0
            if nl >= 20:
                want.append(f"{nl}\t{o}\t{sg}\n")
        assert changed >= 1 and open(b + ".mums").read() == "".join(want)


def test_simt_pfp_last_phrase_suffix_tied_with_a_phrase_end(simt_lib, oracle_bin, tmp_path):
    """Found by tests/fuzz_simt.py: a phrase of this collection ends in '$' / 'A$', so the last phrase's suffixes '$' + padding and
    'A$' + padding share their dictionary group with those shorter strings (padding zero = terminator in the sort key); their rank
    must be looked up inside the group, else a non-member is expanded with a wrapped text position."""
    src = os.path.join(util.ROOT, "tests", "golden", "pfp_last_phrase_tie")
    paths = [os.path.join(src, f"d{i}.fa") for i in range(5)]
    for i, (flags, tile) in enumerate(((["-A"], 0), ([], 300), (["-M", "-n"], 1500), (["-f", "2", "-k", "2", "-l", "8"], 64))):
        os.makedirs(str(tmp_path / f"c{i}"), exist_ok=True)
        both_pfp(simt_lib, oracle_bin, str(tmp_path / f"c{i}"), paths, flags, tile_ranks=tile)


def test_simt_doc_stream_switch(simt_lib, oracle_bin, small, tmp_path, monkeypatch):
    """MMB_DOC_STREAM=1 (experimental): with -M the tiles carry a document-id byte per rank and candidates are checked from it."""
    monkeypatch.setenv("MMB_DOC_STREAM", "1")
    for i, (flags, tile) in enumerate(((["-M", "-n"], 700), (["-M", "-n", "-l", "12"], 100000), (["-M", "-n", "-r"], 900))):
        os.makedirs(str(tmp_path / f"d{i}"), exist_ok=True)
        both_pfp(simt_lib, oracle_bin, str(tmp_path / f"d{i}"), small, flags, tile_ranks=tile)


@pytest.mark.parametrize("flags", [[], ["-k", "-1"], ["-f", "3", "-F", "100"], ["-M", "-n"]])
def test_simt_window_kernel(simt_lib, oracle_bin, small, tmp_path, flags, monkeypatch):
    """The power-of-two window-minimum detector (the default when the smallest interval size exceeds 128) must agree with the chunked
    one on both routes."""
    monkeypatch.setenv("MMB_DETECT", "window")
    for sub in ("direct", "pfp"):
        os.makedirs(str(tmp_path / sub), exist_ok=True)
    both(simt_lib, oracle_bin, str(tmp_path / "direct"), small, flags)
    both_pfp(simt_lib, oracle_bin, str(tmp_path / "pfp"), small, flags, tile_ranks=900)


# ------------------------------------------------------------------------------------------------ prefix-free-parse route (tiled stream)
def both_pfp(simt_lib, oracle_bin, tmp, paths, flags, tile_ranks, no_prune=False):
    o, g = os.path.join(tmp, "o"), os.path.join(tmp, "g")
    util.run_cli(oracle_bin, ["-o", o, *flags, *paths])
    st = api.find_files(paths, g, route="pfp", tile_ranks=tile_ranks, no_prune=no_prune, _lib=simt_lib, **util.flags_to_kwargs(flags))
    assert st["route"] == 2
    assert util.outputs(g) == util.outputs(o), flags
    return st


@pytest.mark.parametrize("name", ["strict", "partial", "mem", "anchor", "arrays", "norc", "mem2", "memF", "bumbl"])
def test_simt_pfp_route_equals_oracle(simt_lib, oracle_bin, small, tmp_path, name):
    """Dictionary + parse suffix arrays, tiled enumeration (SA/LCP per tile with halo), same detector: byte-identical outputs, including the
    .sa/.lcp/.bwt arrays, with tiles so small that intervals straddle tile boundaries."""
    st = both_pfp(simt_lib, oracle_bin, str(tmp_path), small, FLAGSETS[name], tile_ranks=700, no_prune=True)
    assert st["pfp_tiles"] > 10 and st["pfp_distinct_phrases"] < st["pfp_phrases"] and not st["pruned"]
    # the default: segments that cannot hold a match are left out of the stream (same bytes, fewer ranks)
    st = both_pfp(simt_lib, oracle_bin, str(tmp_path), small, FLAGSETS[name], tile_ranks=200)
    if name != "arrays":
        assert st["pruned"] and st["stream_suffixes"] < st["text_symbols"]


@pytest.mark.parametrize("block_order", [0, 2])
def test_simt_pfp_route_repeats(simt_lib, oracle_bin, tmp_path, block_order, monkeypatch):
    """Tandem repeats and low-complexity runs: dictionary suffixes shared by many phrases (multi-member groups beyond one block's
    shared-memory sort), long phrases, phrases of minimum length."""
    rng = np.random.default_rng(17)
    unit = util.random_dna(rng, 37)
    body = util.random_dna(rng, 900)
    docs = [
        "A" * 700 + body + unit * 50 + "ACGT" * 40 + body[:400],
        "A" * 300 + util.mutate(rng, body, 0.01) + unit * 30 + "C" * 150 + body[200:600],
        util.mutate(rng, body, 0.02) + "AC" * 400 + unit * 20 + "A" * 100,
    ]
    p = [util.write_fasta(str(tmp_path / f"d{i}.fa"), [(f"d{i}", s)], width=77) for i, s in enumerate(docs)]
    simt_lib.simt_set_block_order(block_order)
    try:
        for flags in ([], ["-f", "4", "-k", "2", "-l", "15", "-A"], ["-r", "-f", "2"]):
            both_pfp(simt_lib, oracle_bin, str(tmp_path), p, flags, tile_ranks=3000)
        if block_order == 0:
            # the 700-symbol run needs ~33 chunk rounds: with a limit of 12 the dictionary falls back to the doubling engine (nothing pruned)
            monkeypatch.setenv("MMB_CHUNK_ROUNDS", "12")
            st = both_pfp(simt_lib, oracle_bin, str(tmp_path), p, [], tile_ranks=3000)
            assert not st["pruned"] and st["stream_suffixes"] == st["text_symbols"] + 1
    finally:
        simt_lib.simt_set_block_order(0)


def test_simt_pfp_route_edge_inputs(simt_lib, oracle_bin, tmp_path):
    rng = np.random.default_rng(5)
    core = util.random_dna(rng, 700)
    p = []
    p.append(util.write_fasta(str(tmp_path / "multi.fa"), [("c1 desc", core[:250]), ("c2", core[250:480]), ("c3", core[480:])], width=50))
    p.append(util.write_fasta(str(tmp_path / "lower.fasta"), [("x", util.mutate(rng, core, 0.02).lower())], width=61))
    p.append(util.write_fasta(str(tmp_path / "iupac.fna"), [("y", util.mutate(rng, core, 0.01)[:300] + "NNNNNRYKMSWBDHVNU" + core[317:])], width=50, newline="\r\n"))
    for flags in ([], ["-f", "2"], ["-k", "-1", "-l", "10", "-A"]):
        both_pfp(simt_lib, oracle_bin, str(tmp_path), p, flags, tile_ranks=0)


def test_simt_anchor_merge_is_the_fold_not_a_kway_min(simt_lib, ref_dir, tmp_path):
    """MUMs of one partition may overlap.  When the fold suppresses an intermediate candidate (here: a zero threshold at 120), the next
    fold step sees the EARLIER intermediate MUM (start 100, 20 symbols left at 130), not the latest-started source MUMs (90 left) that a
    K-way min/max over the sources would use.  Hand-made partition files, reference anchor_merge (src/merge_candidates.cpp:106-157)."""
    import subprocess
    L = 400

    def part(name, rows, nb):
        with open(tmp_path / f"{name}.mums", "w") as f:
            for ln, offs, strands in rows:
                f.write(f"{ln}\t{','.join(map(str, offs))}\t{','.join(strands)}\n")
        t = np.zeros(L, dtype=np.uint16)
        for k, v in nb.items():
            t[k] = v
        t.tofile(tmp_path / f"{name}.athresh")
        return str(tmp_path / f"{name}.mums")

    a = part("a", [(50, (100, 7), "+-"), (100, (120, 1000), "++")], {100: 5, 130: 5})
    b = part("b", [(200, (90, 300), "+-")], {90: 5, 100: 5, 120: 5, 130: 5})
    c = part("c", [(100, (130, 40), "++")], {130: 5})
    subprocess.run([os.path.join(ref_dir, "anchor_merge"), a, b, c, "-o", str(tmp_path / "om")], check=True, capture_output=True)
    api.merge_files([a, b, c], str(tmp_path / "gm"), _lib=simt_lib)
    got = util.outputs(str(tmp_path / "gm"))
    assert got == util.outputs(str(tmp_path / "om"))
    assert got[".mums"] == b"20\t130,7,440,40\t+,-,-,+\n"          # a K-way min over the latest-started sources would report 90


@pytest.mark.parametrize("flags", [[], ["-k", "-1"], ["-f", "3", "-F", "100"], ["-f", "0", "-k", "2"], ["-b"]])
def test_simt_sharded_collection(simt_lib, oracle_bin, tmp_path, flags):
    """one collection, the suffix order cut into K slices at segment boundaries; every slice reports its own matches and the
    outputs concatenated in slice order are the bytes of the unsharded run (= the oracle's)."""
    paths = util.appendix_a(str(tmp_path / "in"), N=6, L=4000, seed=23, identical_ends=True, snp=0.01, width=70)
    o = str(tmp_path / "o")
    util.run_cli(oracle_bin, ["-o", o, *[f for f in flags if f != "-b"], *paths])
    want = util.outputs(o)
    ext = ".mems" if "-f" in flags else ".mums"
    for k in (2, 3, 7):
        got, rows, streamed = b"", 0, 0
        for i in range(k):
            g = str(tmp_path / f"g{k}_{i}")
            try:
                st = api.find_files(paths, g, shard=(i, k), _lib=simt_lib, **util.flags_to_kwargs(flags))
            except RuntimeError as e:               # -b with an empty slice is an error like an empty result (include/mem_finder.hpp:461-462)
                assert "-b" in flags and "zero MUMs" in str(e)
                continue
            assert st["route"] == 2 and st["pruned"]
            streamed += st["stream_suffixes"]
            if "-b" in flags:
                import mumemto_b200.merge_mums as mm
                l, s, d = mm.read_table(g + ".bumbl")
                got += "".join(f"{l[r]}\t{','.join(map(str, s[r]))}\t{','.join('+' if x else '-' for x in d[r])}\n" for r in range(len(l))).encode()
            else:
                got += util.outputs(g)[ext]
        assert got == want[ext], (flags, k)
    with pytest.raises(Exception, match="sharded"):
        api.find_files(paths, str(tmp_path / "bad"), shard=(0, 2), merge=True, anchor_merge=True, _lib=simt_lib)
    with pytest.raises(Exception, match="shard_index"):
        api.find_files(paths, str(tmp_path / "bad"), shard=(2, 2), _lib=simt_lib)


def test_simt_sharded_collection_repeats(simt_lib, oracle_bin, tmp_path):
    """The same on tandem repeats / low-complexity runs: segments longer than the classifier's window, multi-member groups, many
    occurrences per match; slice bounds fall inside long runs and must move to the next segment start."""
    rng = np.random.default_rng(17)
    unit, body = util.random_dna(rng, 37), util.random_dna(rng, 900)
    docs = ["A" * 700 + body + unit * 50 + "ACGT" * 40 + body[:400],
            "A" * 300 + util.mutate(rng, body, 0.01) + unit * 30 + "C" * 150 + body[200:600],
            util.mutate(rng, body, 0.02) + "AC" * 400 + unit * 20 + "A" * 100]
    p = [util.write_fasta(str(tmp_path / f"d{i}.fa"), [(f"d{i}", s)], width=77) for i, s in enumerate(docs)]
    for flags, ks in (([], (2, 5)), (["-f", "4", "-k", "2", "-l", "15"], (3,))):
        o = str(tmp_path / "o")
        util.run_cli(oracle_bin, ["-o", o, *flags, *p])
        want = util.outputs(o)
        ext = ".mems" if "-f" in flags else ".mums"
        for k in ks:
            got = b""
            for i in range(k):
                g = str(tmp_path / f"g{k}_{i}")
                api.find_files(p, g, shard=(i, k), _lib=simt_lib, **util.flags_to_kwargs(flags))
                got += util.outputs(g)[ext]
            assert got == want[ext], (flags, k)


def test_simt_triggers_scanned_ahead_of_the_parse(simt_lib, oracle_bin, tmp_path, monkeypatch):
    """Host-image loads arrive in upload groups; the text of the finished groups is built AND its trigger windows are scanned while later
    groups are still in flight (engine.cu load_common, pfp.cu PfpTriggers).  Tiny groups here, so that the parse takes over a prefix that
    was scanned ahead in several steps (block-aligned and not) and scans the rest itself; results must not change."""
    monkeypatch.setenv("MMB_UPLOAD_GROUP_BYTES", "3000")
    monkeypatch.setenv("MMB_TRIGGER_AHEAD_MIN", "0")
    paths = util.appendix_a(str(tmp_path / "in"), N=7, L=2900, seed=31, identical_ends=True, snp=0.01, width=70)
    for flags in ([], ["-f", "3", "-F", "50", "-k", "2"], ["-M", "-n"]):
        st = both_pfp(simt_lib, oracle_bin, str(tmp_path), paths, flags, tile_ranks=0)
        assert st["route"] == 2
    monkeypatch.setenv("MMB_UPLOAD_GROUP_BYTES", "9000")        # two files per group; and the direct route must simply drop the masks
    both(simt_lib, oracle_bin, str(tmp_path), paths, [])
