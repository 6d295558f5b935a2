import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _make(*targets, cwd=ROOT):
    r = subprocess.run(["make", "--no-print-directory", *targets], cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"make {' '.join(targets)} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")


@pytest.fixture(scope="session")
def oracle_bin():
    """Our CPU restatement (oracle/oracle.cpp)."""
    _make("port", cwd=os.path.join(ROOT, "oracle"))
    return os.path.join(ROOT, "oracle", "_build", "mumemto_oracle")


@pytest.fixture(scope="session")
def ref_dir():
    """The reference's own sources compiled unmodified (oracle/_ref); built here when /root/reference exists, else the prebuilt copy."""
    d = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir("/root/reference/src"):
        _make("ref", cwd=os.path.join(ROOT, "oracle"))
    if not os.path.exists(os.path.join(d, "mumemto_exec")):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return d


@pytest.fixture(scope="session")
def simt_lib():
    """CPU unit-test build of the kernel sources (tests only)."""
    import ctypes
    from mumemto_b200 import _binding
    _make("simt")
    return _binding.declare(ctypes.CDLL(os.path.join(ROOT, "tests", "simt", "_build", "libmumemto_b200_simt.so")))


@pytest.fixture(scope="session")
def cuda_lib():
    from mumemto_b200 import _binding
    return _binding.load_library()
