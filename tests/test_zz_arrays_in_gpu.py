"""GPU twin of tests/test_arrays_in.py: `-a PREFIX` through the CUDA library and the drop-in executable (runs last: the file name sorts
after the other GPU tests)."""
import os
import subprocess

import pytest

import util
from mumemto_b200 import api

pytestmark = pytest.mark.gpu

REF = os.path.join(util.ROOT, "oracle", "_ref", "mumemto_exec")
EXE = os.path.join(util.ROOT, "mumemto_b200", "bin", "mumemto_exec")


def test_arrays_in_gpu(cuda_lib, oracle_bin, tmp_path):
    paths = util.appendix_a(str(tmp_path / "data"), N=5, L=20000, seed=42, identical_ends=True)
    full = str(tmp_path / "full")
    api.find_files(paths, full, arrays_out=True, _lib=cuda_lib)             # our own -A dump (byte-identical to the reference's, test_t2_golden[arrays])
    for i, flags in enumerate(([], ["-k", "-1"], ["-f", "3", "-F", "100"], ["-M", "-n"], ["-b"])):
        want = str(tmp_path / f"w{i}")
        if os.path.exists(REF):
            subprocess.run([REF, "-o", want, "-a", full, *flags], check=True, capture_output=True)
        else:                                                               # t2's last rank closes nothing: the normal run is the same
            util.run_cli(oracle_bin, ["-o", want, *flags, *paths])
        g, c = str(tmp_path / f"g{i}"), str(tmp_path / f"c{i}")
        st = api.find_arrays(full, g, tile_ranks=(0 if i % 2 else 50000), _lib=cuda_lib, **util.flags_to_kwargs(flags))
        assert st["route"] == 3
        assert util.outputs(g) == util.outputs(want), flags
        r = subprocess.run([EXE, "-o", c, "-a", full, *flags], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        assert util.outputs(c) == util.outputs(want), flags


def test_unbounded_frequency_gpu(cuda_lib, oracle_bin, tmp_path):
    """-f 0 without -F on the GPU: the leftward-walk detector over the whole stream (direct route), two-pass report order."""
    import numpy as np
    rng = np.random.default_rng(17)
    unit, body = util.random_dna(rng, 37), util.random_dna(rng, 3000)
    docs = ["A" * 900 + body + unit * 40 + "ACGT" * 100 + body[:700] + "G" * 450,
            "A" * 300 + util.mutate(rng, body, 0.01) + unit * 26 + "C" * 150 + body[200:900],
            util.mutate(rng, body, 0.02) + "AC" * 270 + unit * 12 + "A" * 100]
    rep = [util.write_fasta(str(tmp_path / f"r{i}.fa"), [(f"r{i}", s)], width=77) for i, s in enumerate(docs)]
    t2 = util.appendix_a(str(tmp_path / "t2"), N=5, L=20000, seed=42, identical_ends=True)
    for name, paths in (("t2", t2), ("rep", rep)):
        for i, flags in enumerate((["-f", "0"], ["-f", "0", "-k", "2", "-l", "12"])):
            o, g = str(tmp_path / f"o_{name}{i}"), str(tmp_path / f"g_{name}{i}")
            util.run_cli(oracle_bin, ["-o", o, *flags, *paths])
            api.find_files(paths, g, _lib=cuda_lib, **util.flags_to_kwargs(flags))
            assert util.outputs(g) == util.outputs(o), (name, flags)


def test_pfp_last_phrase_suffix_tied_with_a_phrase_end_gpu(cuda_lib, oracle_bin, tmp_path):
    """GPU twin of test_simt_pfp_last_phrase_suffix_tied_with_a_phrase_end (a case found by tests/fuzz_simt.py)."""
    src = os.path.join(util.ROOT, "tests", "golden", "pfp_last_phrase_tie")
    paths = [os.path.join(src, f"d{i}.fa") for i in range(5)]
    for i, (flags, tile) in enumerate(((["-A"], 0), ([], 300), (["-M", "-n"], 1500))):
        o, g = str(tmp_path / f"o{i}"), str(tmp_path / f"g{i}")
        util.run_cli(oracle_bin, ["-o", o, *flags, *paths])
        st = api.find_files(paths, g, route="pfp", tile_ranks=tile, _lib=cuda_lib, **util.flags_to_kwargs(flags))
        assert st["route"] == 2
        assert util.outputs(g) == util.outputs(o), flags
