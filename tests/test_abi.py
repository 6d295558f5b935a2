"""The C-ABI shared library loads, exports every symbol the public headers declare, and refuses to work without a CUDA device
(no compute calls here; CPU only)."""
import ctypes
import os
import re
import subprocess

import pytest

import util

LIB = os.path.join(util.ROOT, "mumemto_b200", "lib", "libmumemto_b200.so")


def declared(header, macro):
    text = open(os.path.join(util.ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#define[^\n]*", "", text)
    return sorted(set(re.findall(macro + r"[^;(]*?\b(\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(LIB):
        r = subprocess.run(["make", "-j8", "lib"], cwd=util.ROOT, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    return LIB


def test_every_declared_symbol_is_exported(built_lib):
    exported = set(re.findall(r" T (\w+)", subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout))
    engine = declared("mumemto_b200.h", "MMB_API")
    library = declared("mumemto.h", "MUMEMTO_C_API")
    assert len(engine) >= 30 and len(library) == 15
    missing = [s for s in engine + library if s not in exported]
    assert not missing, missing


def test_binding_signatures_cover_the_header(built_lib):
    from mumemto_b200 import _binding
    assert sorted(_binding._SIGNATURES) == declared("mumemto_b200.h", "MMB_API")


def test_no_device_means_loud_failure(built_lib):
    """In the GPU-less container the product must fail, not fall back."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from mumemto_b200 import Engine, MumemtoError, api
    with pytest.raises(MumemtoError, match="no usable CUDA device"):
        Engine()
    with pytest.raises(MumemtoError, match="no CPU path"):
        api.mum([["ACGTACGTACGT"], ["ACGTACGTACGA"]])
    lib = ctypes.CDLL(built_lib)
    res = ctypes.c_void_p()
    lib.mumemto_last_error.restype = ctypes.c_char_p
    class DocView(ctypes.Structure):
        _fields_ = [("records", ctypes.POINTER(ctypes.c_char_p)), ("num_records", ctypes.c_size_t)]
    recs = (ctypes.c_char_p * 1)(b"ACGTACGT")
    docs = (DocView * 2)(DocView(recs, 1), DocView(recs, 1))
    lib.mumemto_mum.argtypes = [ctypes.POINTER(DocView), ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint8, ctypes.c_size_t, ctypes.c_uint8, ctypes.POINTER(ctypes.c_void_p)]
    assert lib.mumemto_mum(docs, 2, 20, 1, 0, 0, ctypes.byref(res)) == 3          # reference code for "exception"
    assert b"no usable CUDA device" in lib.mumemto_last_error()
    assert lib.mumemto_mum(docs, 2, 20, 1, 0, 0, None) == 1                        # null out_result
    assert lib.mumemto_mum(None, 2, 20, 1, 0, 0, ctypes.byref(res)) == 2           # null docs


def test_cli_binaries_exist_and_reject_bad_input(built_lib, tmp_path):
    exe = os.path.join(util.ROOT, "mumemto_b200", "bin", "mumemto_exec")
    if not os.path.exists(exe):
        r = subprocess.run(["make", "-j8", "all"], cwd=util.ROOT, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    a = util.write_fasta(str(tmp_path / "a.fa"), [("a", "ACGT" * 30)])
    r = subprocess.run([exe, "-o", str(tmp_path / "x"), a], capture_output=True, text=True)
    assert r.returncode == 1 and "Multiple FASTA inputs required" in r.stderr
    r = subprocess.run([exe, "-o", str(tmp_path / "x"), a, str(tmp_path / "missing.fa")], capture_output=True, text=True)
    assert r.returncode == 1 and "not valid" in r.stderr
    open(tmp_path / "b.txt", "w").write(">b\nACGT\n")
    r = subprocess.run([exe, "-o", str(tmp_path / "x"), a, str(tmp_path / "b.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "not a FASTA file" in r.stderr
    r = subprocess.run([exe, "-o", str(tmp_path / "x"), "-f", "-2", a, a], capture_output=True, text=True)
    assert r.returncode == 1


def _core():
    import glob
    if not glob.glob(os.path.join(util.ROOT, "mumemto_b200", "_mumemto_core*.so")):
        r = subprocess.run(["make", "pymodule"], cwd=util.ROOT, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    from mumemto_b200 import _mumemto_core
    return _mumemto_core


def test_pybind_module_has_the_reference_surface(built_lib):
    """python_bindings/src/mumemto_pybind.cpp:93-172: same names, argument names and defaults; without a GPU the call raises."""
    core = _core()
    for name in ("mum", "mem", "MumResult", "MemResult"):
        assert hasattr(core, name)
    for cls in (core.MumResult, core.MemResult):
        for m in ("num_docs", "num_matches", "match_at", "__len__", "__getitem__"):
            assert hasattr(cls, m)
    doc = core.mum.__doc__ + core.mem.__doc__
    for arg in ("sequences", "min_match_len", "use_revcomp", "num_distinct", "use_gsacak", "max_total_freq", "max_doc_freq"):
        assert arg in doc
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match="no CPU path"):
        core.mum([["ACGTACGTACGT"], ["ACGTACGTACGA"]])


CPP_PROGRAM = r"""
#include <cstdio>
#include <mumemto_api.hpp>
int main() {
    std::vector<std::vector<std::string>> docs = {{"ACGTTGCAAGGCTTAACCGGATATCGCGAATTCC", "GGGTTTAAACCC"}, {"TTACGTTGCAAGGCTTAACCGGATATCGCGAATTCCAA"}};
    try {
        mumemto::MumResult r = mumemto::mumemto_mum(docs, 20, true, 0, false);
        std::printf("mums %zu lengths %zu %zu\n", r.matches.size(), r.lengths[0].size(), r.lengths[1][0]);
        for (auto& m : r.matches) std::printf("%u %lld %lld %d %d\n", m.length, (long long)m.offsets[0], (long long)m.offsets[1], m.strands[0], m.strands[1]);
        mumemto::MemResult q = mumemto::mumemto_mem(docs, 20, true, 2, 0, 3);
        std::printf("mems %zu\n", q.matches.size());
    } catch (const std::exception& e) { std::printf("exception: %s\n", e.what()); return 3; }
    return 0;
}
"""


def _build_cpp(tmp_path):
    src = tmp_path / "user.cpp"
    src.write_text(CPP_PROGRAM)
    exe = tmp_path / "user"
    libdir = os.path.dirname(LIB)
    r = subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(util.ROOT, "include"), str(src), "-o", str(exe), "-L" + libdir, "-lmumemto_b200",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return str(exe)


def test_cpp_api_header_compiles_and_fails_loudly_without_a_device(built_lib, tmp_path):
    """include/mumemto_api.hpp = mumemto_library/mumemto_api.hpp:29-57 over the C ABI."""
    exe = _build_cpp(tmp_path)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
def test_pybind_module_and_cpp_api_on_device(built_lib, tmp_path):
    """Same matches through the pybind11 module, the C++ header and the ctypes mirror; views are read-only and keep the result alive."""
    import numpy as np
    from mumemto_b200 import api
    core = _core()
    rng = np.random.default_rng(5)
    base = util.random_dna(rng, 3000)
    docs = [[util.mutate(rng, base, 0.01)[:1700], util.mutate(rng, base, 0.01)[1700:]] for _ in range(4)]
    ref = api.mum(docs, min_match_len=20)
    got = core.mum(docs, min_match_len=20)
    assert len(got) == len(ref) > 0 and got.num_docs() == 4 and got.record_lengths() == [[1700, 1300]] * 4
    for i in (0, len(ref) // 2, -1):
        l, off, st = got[i]
        rl, roff, rst = ref.match_at(i if i >= 0 else len(ref) - 1)
        assert l == rl and (off == roff).all() and (st == rst.astype(bool)).all()
        assert off.dtype == np.int64 and st.dtype == np.bool_ and not off.flags.writeable
    l, off, st = got[0]
    keep = off
    del got, off, st                                             # the view keeps the result object alive
    assert keep.sum() == ref.offsets[0].sum()
    m_ref = api.mem(docs, min_match_len=20, num_distinct=2, max_doc_freq=3)
    m_got = core.mem(docs, min_match_len=20, num_distinct=2, max_doc_freq=3)
    assert len(m_got) == len(m_ref) > 0
    l, off, ids, st = m_got[len(m_got) // 2]
    rl, roff, rids, rst = m_ref.match_at(len(m_ref) // 2)
    assert l == rl and (off == roff).all() and (ids == rids).all() and (st == rst.astype(bool)).all()
    with pytest.raises(IndexError):
        m_got[len(m_got)]
    with pytest.raises(RuntimeError):
        core.mem(docs, max_doc_freq=1)
    out = subprocess.run([_build_cpp(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("mums ") and "lengths 2 38" in out.stdout, out.stdout
