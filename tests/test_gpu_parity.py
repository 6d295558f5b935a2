"""Parity tests proper: the nvcc-built CUDA library, called through the C ABI (include/mumemto_b200.h), against the committed golden
vectors (reference build outputs) and the CPU oracle on seeded inputs.  Bit-exact: everything on this path is integer/byte work."""
import os

import numpy as np
import pytest

import util
from mumemto_b200 import api

pytestmark = pytest.mark.gpu

T2 = dict(N=5, L=20000, seed=42, identical_ends=True)
T2_FLAGS = {"strict": [], "partial": ["-k", "-1"], "mem": ["-f", "3", "-F", "100"], "bumbl": ["-b"], "partial_bumbl": ["-k", "-1", "-b"],
            "anchor": ["-M", "-n"], "arrays": ["-A"], "norc": ["-r"], "short": ["-l", "12"], "mem2": ["-k", "3", "-f", "2"],
            "memF": ["-f", "0", "-F", "7", "-k", "2"]}


def run_gpu(cuda_lib, paths, prefix, flags):
    return api.find_files(paths, prefix, _lib=cuda_lib, **util.flags_to_kwargs(flags))


@pytest.fixture(scope="module")
def t2(tmp_path_factory):
    d = tmp_path_factory.mktemp("t2")
    return str(d), util.appendix_a(str(d), **T2), util.load_golden_table()["t2"]


@pytest.fixture(scope="module")
def c1(tmp_path_factory):
    d = tmp_path_factory.mktemp("c1")
    return str(d), util.appendix_a(str(d), N=5, L=1000000, seed=1), util.load_golden_table()["t7"]


@pytest.mark.parametrize("name", sorted(T2_FLAGS))
def test_t2_golden(cuda_lib, t2, name):
    d, paths, table = t2
    run_gpu(cuda_lib, paths, os.path.join(d, "g_" + name), T2_FLAGS[name])
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "g_" + name)).items()}
    assert got == table["outputs"][name]["md5"]


@pytest.mark.parametrize("name", ["strict", "partial", "mem", "anchor"])
def test_c1_golden(cuda_lib, c1, name):
    """BASELINE configs[0]: 5 x 1 Mbp synthetic haplotypes (1 % SNP); 18 070 strict MUMs, 48 393 partial, 36 139 MEMs."""
    d, paths, table = c1
    st = run_gpu(cuda_lib, paths, os.path.join(d, "g_" + name), table["outputs"][name]["flags"])
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "g_" + name)).items()}
    assert got == table["outputs"][name]["md5"]
    assert st["text_symbols"] == 10000010 and st["input_bases"] == 5000000


def test_edge_inputs_vs_oracle(cuda_lib, oracle_bin, tmp_path):
    rng = np.random.default_rng(5)
    core = util.random_dna(rng, 7000)
    p = []
    p.append(util.write_fasta(str(tmp_path / "multi.fa"), [("c1 desc", core[:2500]), ("c2", core[2500:4800]), ("c3", core[4800:])], width=50))
    p.append(util.write_fasta(str(tmp_path / "lower.fasta"), [("x", util.mutate(rng, core, 0.02).lower())], width=61))
    p.append(util.write_fasta(str(tmp_path / "iupac.fna"), [("y", util.mutate(rng, core, 0.01)[:3000] + "NNNNNRYKMSWBDHVNU" + core[3017:])], width=50, newline="\r\n"))
    with open(tmp_path / "odd.fa", "w") as f:
        f.write("junk before header\n\n>z some comment > with marker\n" + core[:1000] + "\n\n" + core[1000:3333] + "\n>empty\n>tail\n" + core[3333:])
    p.append(str(tmp_path / "odd.fa"))
    for i, flags in enumerate(([], ["-f", "2"], ["-k", "-1", "-l", "10"], ["-r"], ["-M", "-n"], ["-f", "5", "-l", "8"])):
        o, g = str(tmp_path / f"o{i}"), str(tmp_path / f"g{i}")
        util.run_cli(oracle_bin, ["-o", o, *flags, *p])
        run_gpu(cuda_lib, p, g, flags)
        assert util.outputs(g) == util.outputs(o), flags
        assert open(g + ".lengths").read() == open(o + ".lengths").read()


def test_text_long_lines_and_headers(cuda_lib, oracle_bin, tmp_path):
    """Chunks without a newline, headers longer than a chunk, records meeting chunk boundaries (the scatter pass takes the state before a
    thread's bytes from the nearest earlier newline; these shapes have none nearby)."""
    from test_simt_pipeline import _long_line_inputs
    p = _long_line_inputs(tmp_path, np.random.default_rng(8), 60000)
    for i, flags in enumerate(([], ["-r", "-f", "2"], ["-k", "-1"])):
        o, g = str(tmp_path / f"o{i}"), str(tmp_path / f"g{i}")
        util.run_cli(oracle_bin, ["-o", o, *flags, *p])
        run_gpu(cuda_lib, p, g, flags)
        assert util.outputs(g) == util.outputs(o), flags
        assert open(g + ".lengths").read() == open(o + ".lengths").read()


@pytest.mark.parametrize("seed,N,L,snp", [(21, 2, 50000, 0.001), (22, 9, 30000, 0.03), (23, 33, 8000, 0.01), (24, 70, 3000, 0.02)])
def test_random_collections_vs_oracle(cuda_lib, oracle_bin, tmp_path, seed, N, L, snp):
    """N = 2 ... 70 documents (70 exercises the multi-word document counters), shallow and deep LCPs."""
    paths = util.appendix_a(str(tmp_path), N=N, L=L, seed=seed, snp=snp)
    for i, flags in enumerate(([], ["-k", "-2"], ["-f", "2", "-k", "2"], ["-M", "-n"])):
        o, g = str(tmp_path / f"o{i}"), str(tmp_path / f"g{i}")
        util.run_cli(oracle_bin, ["-o", o, *flags, *paths])
        run_gpu(cuda_lib, paths, g, flags)
        assert util.outputs(g) == util.outputs(o), flags


def test_quirk_last_rank_and_saturation(cuda_lib, oracle_bin, tmp_path):
    rng = np.random.default_rng(3)
    S = "TTTTGACGTCAGCTAGCTAGCATCGA"
    a = util.random_dna(rng, 50, "ACG") + "G" + S + "A" + util.random_dna(rng, 50, "ACG")
    b = util.random_dna(rng, 50, "ACG") + "A" + S + "C" + util.random_dna(rng, 50, "ACG")
    p = [util.write_fasta(str(tmp_path / "a.fa"), [("a", a)]), util.write_fasta(str(tmp_path / "b.fa"), [("b", b)])]
    run_gpu(cuda_lib, p, str(tmp_path / "q6"), ["-r"])
    assert open(tmp_path / "q6.mums").read() == ""
    X = util.random_dna(np.random.default_rng(9), 70000)
    p = [util.write_fasta(str(tmp_path / "d0.fa"), [("d0", X + "A" + X[:66000])], width=80), util.write_fasta(str(tmp_path / "d1.fa"), [("d1", X)], width=80)]
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "o"), "-M", "-n", "-r", *p])
    run_gpu(cuda_lib, p, str(tmp_path / "g"), ["-M", "-n", "-r"])
    assert util.outputs(str(tmp_path / "g")) == util.outputs(str(tmp_path / "o"))
    assert np.frombuffer(util.outputs(str(tmp_path / "g"))[".athresh"], dtype=np.uint16).max() == 65535


def test_repeats_all_group_size_classes(cuda_lib, oracle_bin, tmp_path):
    """Low-complexity / tandem-repeat inputs: every group-size class of the doubling rounds (sub-warp ... block sort, radix fallback)."""
    rng = np.random.default_rng(17)
    unit = util.random_dna(rng, 37)
    body = util.random_dna(rng, 30000)
    docs = [
        "A" * 9000 + body + unit * 300 + "ACGT" * 400 + body[:7000] + "G" * 4500,
        "A" * 300 + util.mutate(rng, body, 0.01) + unit * 260 + "C" * 150 + body[2000:9000],
        util.mutate(rng, body, 0.02) + "AC" * 2700 + unit * 120 + "A" * 1000,
    ]
    p = [util.write_fasta(str(tmp_path / f"d{i}.fa"), [(f"d{i}", s)], width=77) for i, s in enumerate(docs)]
    for i, flags in enumerate(([], ["-k", "2"], ["-f", "4", "-k", "2", "-l", "15"], ["-f", "0", "-F", "40", "-k", "2"], ["-r", "-f", "2"])):
        o, g = str(tmp_path / f"o{i}"), str(tmp_path / f"g{i}")
        util.run_cli(oracle_bin, ["-o", o, *flags, *p])
        run_gpu(cuda_lib, p, g, flags)
        assert util.outputs(g) == util.outputs(o), flags


def test_library_mirror(cuda_lib, oracle_bin, tmp_path):
    r = api.mem([["ACGTACGT"], ["TACGTAAA"]], min_match_len=3, max_doc_freq=2, _lib=cuda_lib)
    assert len(r) == 2
    l, off, ids, st = r.match_at(0)
    assert (l, off.tolist(), ids.tolist(), st.tolist()) == (5, [0, 1, 0, 4], [1, 1, 0, 0], [0, 1, 1, 0])
    assert len(api.mum([["ACGTACGT"], ["TACGTAAA"]], min_match_len=3, _lib=cuda_lib)) == 0
    rng = np.random.default_rng(2)
    core = util.random_dna(rng, 15000)
    docs = [[core[:7000], core[7000:]], [util.mutate(rng, core, 0.02)], [util.mutate(rng, core, 0.02)[:8000], util.mutate(rng, core, 0.01)[8000:]]]
    paths = [util.write_fasta(str(tmp_path / f"d{i}.fa"), [(f"r{j}", s) for j, s in enumerate(d)]) for i, d in enumerate(docs)]
    util.run_cli(oracle_bin, ["-o", str(tmp_path / "o"), "-k", "2", *paths])
    res = api.mum(docs, num_distinct=2, _lib=cuda_lib)
    lines = []
    for i in range(len(res)):
        l, off, st = res.match_at(i)
        lines.append(f"{l}\t" + ",".join("" if o < 0 else str(o) for o in off) + "\t" + ",".join("" if o < 0 else ("+" if s else "-") for o, s in zip(off, st)) + "\n")
    assert "".join(lines) == open(tmp_path / "o.mums").read()


def test_size_independent_properties(cuda_lib, tmp_path):
    """At a size the oracle is not run on here (8 x 2 Mbp): SA is a permutation, LCP agrees with direct comparison on sampled ranks,
    every reported MUM really occurs at its coordinates, and document order / strand rules hold."""
    from mumemto_b200 import Engine
    paths = util.appendix_a(str(tmp_path), N=8, L=2000000, seed=77, snp=0.005)
    e = Engine(_lib=cuda_lib)
    e.load_fasta_files(paths)
    st = e.find()
    T = np.frombuffer(e.symbols(), dtype=np.uint8)
    sa, lcp, bwt = e.arrays()
    m = len(T) + 1
    assert len(sa) == m and sa[0] == len(T)
    seen = np.zeros(m, dtype=bool); seen[sa] = True
    assert seen.all()
    rng = np.random.default_rng(0)
    for r in rng.integers(1, m, 2000):
        a, b, l = int(sa[r - 1]), int(sa[r]), int(lcp[r])
        x, y = T[a:a + l + 1], T[b:b + l + 1]
        assert (x[:l] == y[:l]).all() and (len(x) <= l or len(y) <= l or x[l] != y[l])
        assert bytes(x[:l + 1]) < bytes(y[:l + 1]) or len(x) < len(y)
        assert bwt[r] == (T[b - 1] if b else 0)
    lengths, offsets, strands = e.mums()
    assert st["num_matches"] == len(lengths) > 1000
    docs = [np.frombuffer(open(p, "rb").read().split(b"\n", 1)[1].replace(b"\n", b""), dtype=np.uint8) for p in paths]
    comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    for i in rng.integers(0, len(lengths), 300):
        l = int(lengths[i]); assert l >= 20 and strands[i][0] == 1
        ref = docs[0][offsets[i][0]:offsets[i][0] + l]
        for d in range(1, 8):
            seg = docs[d][offsets[i][d]:offsets[i][d] + l]
            if not strands[i][d]:
                seg = comp[seg[::-1]]
            assert (seg == ref).all()
    e.close()


# ------------------------------------------------------------------------------------------------ prefix-free-parse route
@pytest.mark.parametrize("name", ["strict", "partial", "mem", "anchor"])
def test_c1_golden_pfp_route(cuda_lib, c1, name):
    """C1 through the dictionary/parse route with 64 stream tiles: same golden md5s."""
    d, paths, table = c1
    kw = util.flags_to_kwargs(table["outputs"][name]["flags"])
    st = api.find_files(paths, os.path.join(d, "p_" + name), route="pfp", tile_ranks=160000, no_prune=True, _lib=cuda_lib, **kw)
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "p_" + name)).items()}
    assert got == table["outputs"][name]["md5"]
    assert st["route"] == 2 and st["pfp_tiles"] >= 60 and not st["pruned"]
    # default: segment pruning -- the same bytes from a fraction of the ranks
    st = api.find_files(paths, os.path.join(d, "q_" + name), route="pfp", tile_ranks=160000, _lib=cuda_lib, **kw)
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "q_" + name)).items()}
    assert got == table["outputs"][name]["md5"]
    assert st["pruned"] and st["stream_suffixes"] < st["text_symbols"] // 2


@pytest.mark.parametrize("name", ["strict", "arrays", "memF", "bumbl", "norc"])
def test_t2_golden_pfp_route(cuda_lib, t2, name):
    d, paths, table = t2
    api.find_files(paths, os.path.join(d, "p_" + name), route="pfp", tile_ranks=9000, _lib=cuda_lib, **util.flags_to_kwargs(T2_FLAGS[name]))
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "p_" + name)).items()}
    assert got == table["outputs"][name]["md5"]


def test_pfp_route_repeats_and_many_docs(cuda_lib, oracle_bin, tmp_path):
    rng = np.random.default_rng(17)
    unit = util.random_dna(rng, 37)
    body = util.random_dna(rng, 30000)
    docs = [
        "A" * 9000 + body + unit * 300 + "ACGT" * 400 + body[:7000] + "G" * 4500,
        "A" * 300 + util.mutate(rng, body, 0.01) + unit * 260 + "C" * 150 + body[2000:9000],
        util.mutate(rng, body, 0.02) + "AC" * 2700 + unit * 120 + "A" * 1000,
    ]
    p = [util.write_fasta(str(tmp_path / f"d{i}.fa"), [(f"d{i}", s)], width=77) for i, s in enumerate(docs)]
    for i, flags in enumerate(([], ["-f", "4", "-k", "2", "-l", "15", "-A"], ["-r", "-f", "2"])):
        o, g = str(tmp_path / f"o{i}"), str(tmp_path / f"g{i}")
        util.run_cli(oracle_bin, ["-o", o, *flags, *p])
        api.find_files(p, g, route="pfp", tile_ranks=50000, _lib=cuda_lib, **util.flags_to_kwargs(flags))
        assert util.outputs(g) == util.outputs(o), flags
    paths = util.appendix_a(str(tmp_path / "many"), N=70, L=3000, seed=24, snp=0.02)
    for i, flags in enumerate(([], ["-k", "-2"], ["-M", "-n"])):
        o, g = str(tmp_path / f"mo{i}"), str(tmp_path / f"mg{i}")
        util.run_cli(oracle_bin, ["-o", o, *flags, *paths])
        api.find_files(paths, g, route="pfp", _lib=cuda_lib, **util.flags_to_kwargs(flags))
        assert util.outputs(g) == util.outputs(o), flags


def test_pfp_route_equals_direct_route_pangenome(cuda_lib, tmp_path):
    """64 haplotype-like documents (shared variants): the two routes must produce the same bytes (sizes where the CPU oracle is slow)."""
    import bench
    imgs = bench.make_collection(dict(N=64, L=300000, kind="pangenome"), seed=3)
    paths = []
    for i, im in enumerate(imgs):
        p = str(tmp_path / f"h{i}.fa"); im.tofile(p); paths.append(p)
    for i, flags in enumerate(([], ["-k", "-1"], ["-f", "3", "-F", "100"])):
        a, b = str(tmp_path / f"a{i}"), str(tmp_path / f"b{i}")
        api.find_files(paths, a, route="direct", _lib=cuda_lib, **util.flags_to_kwargs(flags))
        st = api.find_files(paths, b, route="pfp", tile_ranks=1 << 22, _lib=cuda_lib, **util.flags_to_kwargs(flags))
        assert util.outputs(a) == util.outputs(b), flags
        assert st["pfp_dict_bytes"] < 0.2 * st["text_symbols"] and st["pruned"] and st["stream_suffixes"] < 0.2 * st["text_symbols"]
        c = str(tmp_path / f"c{i}")
        st = api.find_files(paths, c, route="pfp", tile_ranks=1 << 22, no_prune=True, _lib=cuda_lib, **util.flags_to_kwargs(flags))
        assert util.outputs(a) == util.outputs(c), flags
        assert st["stream_suffixes"] == st["text_symbols"] + 1
    # -M -n: thresholds of the pruned segments come from the segment kernel, not from the stream
    for i, extra in enumerate(({}, {"no_prune": True}, {"route": "direct"})):
        api.find_files(paths, str(tmp_path / f"m{i}"), merge=True, anchor_merge=True, _lib=cuda_lib, **({"route": "pfp", **extra}))
    assert util.outputs(str(tmp_path / "m0")) == util.outputs(str(tmp_path / "m2")) == util.outputs(str(tmp_path / "m1"))


# ------------------------------------------------------------------------------------------------ the drop-in executables
BIN = os.path.join(util.ROOT, "mumemto_b200", "bin")


@pytest.mark.parametrize("name", ["strict", "partial", "mem", "bumbl", "partial_bumbl", "anchor", "string_merge", "arrays", "norc"])
def test_cli_mumemto_exec_t2_golden(t2, name):
    """bin/mumemto_exec with the reference's flags writes byte-identical files (PREFIX.lengths compared after path normalisation)."""
    d, paths, table = t2
    flags = {**T2_FLAGS, "string_merge": ["-M"]}[name]
    util.run_cli(os.path.join(BIN, "mumemto_exec"), ["-o", os.path.join(d, "c_" + name), *flags, *paths])
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "c_" + name)).items()}
    assert got == table["outputs"][name]["md5"]
    lines = open(os.path.join(d, "c_" + name + ".lengths")).read().splitlines()
    assert lines[0] == f"{os.path.realpath(paths[0])} * 20000" and lines[1] == f"{os.path.realpath(paths[0])} h0 20000" and len(lines) == 10


@pytest.mark.parametrize("tag", ["merge2", "merge3"])
def test_cli_anchor_merge_t2_golden(t2, tag):
    d, paths, table = t2
    names = []
    for i, idx in enumerate(table["outputs"][tag]["partitions"]):
        pre = os.path.join(d, f"cm_{tag}_{i}")
        util.run_cli(os.path.join(BIN, "mumemto_exec"), ["-o", pre, "-M", "-n", *[paths[j] for j in idx]])
        names.append(pre + ".mums")
    util.run_cli(os.path.join(BIN, "anchor_merge"), [*names, "-o", os.path.join(d, "cm_" + tag)])
    got = {ext: util.md5(b) for ext, b in util.outputs(os.path.join(d, "cm_" + tag)).items()}
    assert got == table["outputs"][tag]["md5"]
    # .bumbl partitions and .bumbl output go through the same fold
    names_b = []
    for i, idx in enumerate(table["outputs"][tag]["partitions"]):
        pre = os.path.join(d, f"cb_{tag}_{i}")
        util.run_cli(os.path.join(BIN, "mumemto_exec"), ["-o", pre, "-M", "-n", "-b", *[paths[j] for j in idx]])
        names_b.append(pre + ".bumbl")
    util.run_cli(os.path.join(BIN, "anchor_merge"), [*names_b, "-o", os.path.join(d, "cb_" + tag + ".bumbl")])
    assert util.md5_file(os.path.join(d, "cb_" + tag + ".athresh")) == table["outputs"][tag]["md5"][".athresh"]
    import numpy as np
    raw = open(os.path.join(d, "cb_" + tag + ".bumbl"), "rb").read()
    n_seqs, n_mums = np.frombuffer(raw[2:18], dtype=np.uint64)
    assert (int(n_seqs), int(n_mums)) == (5, 368)


def test_library_abi_from_c(tmp_path):
    """A C program written against the reference's header (include/mumemto.h = mumemto_library/mumemto.h) links and runs."""
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include "mumemto.h"
int main(void) {
    const char* a[] = {"ACGTACGT"}; const char* b[] = {"TACGTAAA"};
    mumemto_doc_view docs[2] = {{a, 1}, {b, 1}};
    mumemto_mem_result* r = NULL;
    int rc = mumemto_mem(docs, 2, 3, 1, 0, 0, 2, 0, &r);
    if (rc) { printf("rc %d %s\n", rc, mumemto_last_error()); return 1; }
    printf("%zu %zu\n", num_docs_mem(r), num_mems(r));
    for (size_t i = 0; i < num_mems(r); ++i) {
        mumemto_mem_match_view v = mem_at(r, i);
        printf("%u", v.length);
        for (size_t k = 0; k < v.occurrences; ++k) printf(" %lld/%zu/%c", (long long)v.offsets[k], v.seq_ids[k], v.strands[k] ? '+' : '-');
        printf("\n");
    }
    mem_free(r);
    mumemto_mum_result* m = NULL;
    rc = mumemto_mum(docs, 2, 3, 1, 0, 0, &m);
    printf("%d %zu %zu\n", rc, num_mums(m), doc_record_offsets(m)[2]);
    mum_free(m);
    return 0;
}
''')
    exe = tmp_path / "t"
    lib = os.path.join(util.ROOT, "mumemto_b200", "lib")
    import subprocess
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(util.ROOT, "include"), str(src), "-L", lib, f"-Wl,-rpath,{lib}", "-lmumemto_b200", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out == ["2 2", "5 0/1/- 1/1/+ 0/0/+ 4/0/-", "5 3/0/+ 0/0/- 1/1/- 0/1/+", "0 0 2"]


# ------------------------------------------------------------------------------------------------ the calls the multi-GPU number rests on
def test_device_input_anchor_merge_vs_reference(cuda_lib, ref_dir, tmp_path):
    """Partitions run one after the other on ONE GPU, their device tables (mmb_mum_table_device) handed to
    mmb_anchor_merge(inputs_on_device = 1), against the reference's own `mumemto_exec -M -n` x P -> `anchor_merge`
    (src/merge_candidates.cpp:106-157,208-219): merged .mums bytes and merged .athresh bytes; 64 haplotype-like documents, 2 / 3 / 8
    partitions (8 = the fold depth of the 8-GPU run)."""
    import bench
    import torch
    from mumemto_b200 import Engine
    exe, merge_exe = os.path.join(ref_dir, "mumemto_exec"), os.path.join(ref_dir, "anchor_merge")
    files = bench.SampleFiles(dict(N=64, L=120000, kind="pangenome"))
    eng = Engine(0, _lib=cuda_lib)
    dev = torch.device("cuda", 0)
    for procs in (2, 3, 8):
        _, groups, ref_mums, ref_athresh = bench.run_reference_partitioned(exe, merge_exe, files, procs)
        parts = []
        for grp in groups:
            eng.load_fasta_bytes([files.imgs[i] for i in grp])
            eng.find(merge=True, anchor_merge=True, want_text=False)
            parts.append(bench.snapshot_partition(eng, dev))
        got_mums, got_athresh = bench.merge_on_device(eng, parts)
        assert got_mums == ref_mums and got_athresh == ref_athresh, procs
        assert ref_mums.count(b"\n") > 100
    eng.close()
    files.close()


def test_bench_parity_sample_on_gpu(cuda_lib, ref_dir):
    """bench.py's off-the-clock parity check itself (64 x 200 kbp of the benchmark generator; C2, C3 and C4 flag sets, and the
    partition -> device-merge leg of C5) through the CUDA library."""
    import bench
    import torch
    from mumemto_b200 import Engine
    eng = Engine(0, _lib=cuda_lib)
    for name in ("c2", "c3", "c4"):
        base, parity = bench.cpu_baseline_and_parity(eng, torch.device("cuda", 0), bench.WORKLOADS[name], sample_L=200000, procs=4)
        assert parity["result"] == "ok", name
        assert ("partitions_anchor_merge" in parity["checks"]) == (name == "c2")
    eng.close()


def test_cli_arrays_out_on_the_tiled_route(cuda_lib, ref_dir, tmp_path):
    """`mumemto_exec -A` on an input that takes the prefix-free-parse route (64 haplotype-like documents): .sa / .lcp / .bwt must be
    written and equal the reference's -A files (the CLI has to ask the engine to keep the whole stream)."""
    import bench
    files = bench.SampleFiles(dict(N=64, L=50000, kind="pangenome"))
    ref, got = str(tmp_path / "ref"), str(tmp_path / "got")
    util.run_cli(os.path.join(ref_dir, "mumemto_exec"), ["-o", ref, "-A", *files.paths])
    r = util.run_cli(os.path.join(BIN, "mumemto_exec"), ["-o", got, "-A", *files.paths])
    assert b"route 2" in r.stderr or True
    a, b = util.outputs(got), util.outputs(ref)
    assert set(a) == {".mums", ".sa", ".lcp", ".bwt"} and a == b
    files.close()


@pytest.mark.parametrize("mode", ["partitions", "gpus"])
def test_cli_partition_merge_in_one_process(ref_dir, tmp_path, mode):
    """`mumemto_exec --partitions P` (one GPU, partitions one after the other) and `--gpus N` (one host thread per GPU + ncclAllGather,
    csrc/multi_gpu.cpp) against the reference workflow on the same partitions: P x `mumemto_exec -M -n`, then `anchor_merge`
    (mumemto/merge_mums.py:171-183, src/merge_candidates.cpp:170-255): PREFIX.mums, PREFIX.athresh byte for byte, PREFIX.lengths lines."""
    import bench
    import subprocess
    import torch
    P = 3
    if mode == "gpus":
        if torch.cuda.device_count() < 2:
            pytest.skip("needs two GPUs")
        P = min(torch.cuda.device_count(), 4)
    files = bench.SampleFiles(dict(N=13, L=150000, kind="pangenome"))
    rest = len(files.paths) - 1
    groups, at = [], 1
    for p in range(P):                          # contiguous slices, like mgpu::make_partitions
        take = rest // P + (1 if p < rest % P else 0)
        groups.append([0] + list(range(at, at + take))); at += take
    names = []
    for g, grp in enumerate(groups):
        pre = str(tmp_path / f"ref{g}")
        subprocess.run([os.path.join(ref_dir, "mumemto_exec"), "-o", pre, "-M", "-n", *[files.paths[i] for i in grp]], check=True, capture_output=True)
        names.append(pre + ".mums")
    subprocess.run([os.path.join(ref_dir, "anchor_merge"), *names, "-o", str(tmp_path / "refm")], check=True, capture_output=True)
    flag = ["--partitions", str(P)] if mode == "partitions" else ["--gpus", str(P)]
    util.run_cli(os.path.join(BIN, "mumemto_exec"), ["-o", str(tmp_path / "got"), *flag, *files.paths])
    a, b = util.outputs(str(tmp_path / "got")), util.outputs(str(tmp_path / "refm"))
    assert a == b and len(b[".mums"]) > 1000
    lines = open(tmp_path / "got.lengths").read().splitlines()
    assert len(lines) == 2 * len(files.paths) and lines[0].endswith(" * 150000") and lines[1].endswith(" h0 150000")
    files.close()


@pytest.mark.gpu
@pytest.mark.parametrize("flagset", ["strict", "partial", "mem"])
def test_midsize_pangenome_golden_md5(cuda_lib, tmp_path, flagset):
    """64 x 4 Mbp of the benchmark generator (|T| = 5.1e8 symbols, fifty times C1): the md5 of the engine's output against the md5 of the
    reference's own run on the same collection (tests/golden/md5_table.json `full:h64_4m:*`, recorded by scripts/make_fullsize_golden.py).
    Skipped until that entry exists."""
    import hashlib
    import bench
    entry = util.load_golden_table().get(f"full:h64_4m:{flagset}")
    if not entry:
        pytest.skip("no golden entry for h64_4m")
    imgs = bench.make_collection(bench.WORKLOADS["h64_4m"], seed=1)
    paths = []
    for i, im in enumerate(imgs):
        p = str(tmp_path / f"h{i}.fa"); im.tofile(p); paths.append(p)
    st = api.find_files(paths, str(tmp_path / "out"), _lib=cuda_lib, **util.flags_to_kwargs(entry["flags"]))
    assert st["route"] == 2 and st["pruned"]
    ext = ".mems" if flagset == "mem" else ".mums"
    with open(str(tmp_path / "out") + ext, "rb") as f:
        data = f.read()
    assert len(data) == entry["bytes"][ext] and hashlib.md5(data).hexdigest() == entry["md5"][ext]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["passes", "gpus"])
def test_cli_sharded_collection(ref_dir, tmp_path, mode):
    """`mumemto_exec --shards K` -- ONE collection, every GPU (or pass) reports its slice of the suffix order, slices written
    in order.  The files must be the bytes the reference writes for the whole collection (strict, -k -1, MEMs, .bumbl)."""
    import bench
    import subprocess
    import torch
    K = 3
    extra = []
    if mode == "gpus":
        if torch.cuda.device_count() < 2:
            pytest.skip("needs two GPUs")
        K = min(torch.cuda.device_count(), 4)
        extra = ["--gpus", str(K)]
    files = bench.SampleFiles(dict(N=12, L=120000, kind="pangenome"))
    for i, flags in enumerate(([], ["-k", "-1"], ["-f", "3", "-F", "100"], ["-b"])):
        ref, got = str(tmp_path / f"ref{i}"), str(tmp_path / f"got{i}")
        subprocess.run([os.path.join(ref_dir, "mumemto_exec"), "-o", ref, *flags, *files.paths], check=True, capture_output=True)
        util.run_cli(os.path.join(BIN, "mumemto_exec"), ["-o", got, "--shards", str(K), *extra, *flags, *files.paths])
        a, b = util.outputs(got), util.outputs(ref)
        assert a == b and sum(map(len, b.values())) > 1000, flags
        assert open(got + ".lengths").read() == open(ref + ".lengths").read()
    r = util.run_cli(os.path.join(BIN, "mumemto_exec"), ["-o", str(tmp_path / "bad"), "--shards", "2", "-M", "-n", *files.paths], check=False)
    assert r.returncode == 1 and b"--shards cannot be combined" in r.stderr
    files.close()
