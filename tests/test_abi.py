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
