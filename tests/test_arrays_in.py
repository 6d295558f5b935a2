"""`-a PREFIX`: detection over pre-computed SA / LCP / BWT files (reference src/pfp_mum.cpp:96-110, include/read_arrays.hpp:66-84,
RefBuilder(prefix) src/ref_builder.cpp:140-168) -- SURVEY.md section 8(f).1.  Checker: the reference's own `-a` run (oracle/_ref) on the
arrays the reference's own `-A` run dumped.  CPU: the SIMT unit-test build of the engine."""
import os
import subprocess

import numpy as np
import pytest

import util
from mumemto_b200 import api

REF = os.path.join(util.ROOT, "oracle", "_ref", "mumemto_exec")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref (the reference's sources compiled unmodified) is not built")

FLAGSETS = ([], ["-k", "-1"], ["-f", "3", "-F", "100"], ["-M", "-n"], ["-b"], ["-l", "12"], ["-f", "2", "-k", "2"])


def ref_arrays(tmp, paths, extra=()):
    full = os.path.join(tmp, "full")
    subprocess.run([REF, "-o", full, "-A", *extra, *paths], check=True, capture_output=True)
    return full


def check_all(lib, tmp, full, extra, tile_ranks):
    for i, flags in enumerate(FLAGSETS):
        r, g = os.path.join(tmp, f"ref{i}"), os.path.join(tmp, f"got{i}")
        subprocess.run([REF, "-o", r, "-a", full, *extra, *flags], check=True, capture_output=True)
        st = api.find_arrays(full, g, tile_ranks=tile_ranks, _lib=lib, **util.flags_to_kwargs([*extra, *flags]))
        assert st["route"] == 3
        assert util.outputs(g) == util.outputs(r), flags
        assert not os.path.exists(g + ".lengths")           # RefBuilder(prefix) never writes one


@needs_ref
def test_simt_arrays_in_t2(simt_lib, tmp_path):
    paths = util.appendix_a(str(tmp_path / "data"), N=5, L=20000, seed=42, identical_ends=True)
    check_all(simt_lib, str(tmp_path), ref_arrays(str(tmp_path), paths), [], tile_ranks=7000)


@needs_ref
def test_simt_arrays_in_no_revcomp_and_one_tile(simt_lib, tmp_path):
    paths = util.appendix_a(str(tmp_path / "data"), N=3, L=6000, seed=9, snp=0.02)
    check_all(simt_lib, str(tmp_path), ref_arrays(str(tmp_path), paths, ["-r"]), ["-r"], tile_ranks=0)


@needs_ref
def test_simt_arrays_in_stops_one_rank_early(simt_lib, tmp_path):
    """The reference's reader walks |T| of the |T|+1 entries, so an interval that the LAST rank closes is reported by the normal run
    but not by -a; ours must do the same (built from the t8 construction of SURVEY.md Appendix A: the repeat is made the
    second-to-last group of the suffix order by one larger suffix, 'TTTTT', behind it)."""
    rng = np.random.default_rng(3)
    S = "TTTTGACGTCAGCTAGCTAGCATCGA"
    def doc(a, b, tail):
        return "".join(rng.choice(list("ACG"), 50)) + a + S + b + "".join(rng.choice(list("ACG"), 50)) + tail
    p = [util.write_fasta(str(tmp_path / "a.fa"), [("a", doc("G", "A", "TTTTT"))]), util.write_fasta(str(tmp_path / "b.fa"), [("b", doc("A", "C", ""))])]
    full = ref_arrays(str(tmp_path), p, ["-r"])
    normal = util.outputs(full)
    r, g = str(tmp_path / "ref"), str(tmp_path / "got")
    subprocess.run([REF, "-o", r, "-a", full, "-r"], check=True, capture_output=True)
    api.find_arrays(full, g, revcomp=False, _lib=simt_lib)
    assert util.outputs(g) == util.outputs(r)
    assert normal[".mums"] == b"26\t51,51\t+,+\n" and util.outputs(r)[".mums"] == b""


def test_simt_arrays_in_round_trip_of_our_own_dump(simt_lib, oracle_bin, tmp_path):
    """Arrays dumped by our engine (-A) fed back through -a reproduce the oracle's normal output on inputs whose last rank closes
    nothing (no dependence on oracle/_ref)."""
    paths = util.appendix_a(str(tmp_path / "data"), N=4, L=5000, seed=21, snp=0.02)
    full = str(tmp_path / "full")
    api.find_files(paths, full, arrays_out=True, _lib=simt_lib)
    for i, flags in enumerate(([], ["-k", "-1"], ["-f", "2", "-k", "2"])):
        o, g = str(tmp_path / f"o{i}"), str(tmp_path / f"g{i}")
        util.run_cli(oracle_bin, ["-o", o, *flags, *paths])
        api.find_arrays(full, g, tile_ranks=3000, _lib=simt_lib, **util.flags_to_kwargs(flags))
        got, want = util.outputs(g), util.outputs(o)
        assert got == want, flags


def test_simt_arrays_in_errors(simt_lib, tmp_path):
    from mumemto_b200._binding import Engine, MumemtoError
    e = Engine(_lib=simt_lib)
    z = np.zeros(64, dtype=np.uint8)
    with pytest.raises(MumemtoError):
        e.find_from_arrays([10], z, z, z, 5)                 # one document
    with pytest.raises(MumemtoError):
        e.find_from_arrays([3, 3], z, z, z, 100)             # more entries than |T| + 1
    with pytest.raises(MumemtoError):
        e.find()                                             # no text loaded
    e.close()
