import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """The CPU suite (`-m "not gpu"`) runs its tests on several worker processes (pytest-xdist) so that it finishes within a few minutes:
    the SIMT unit-test build executes every kernel thread by thread.  Everything the tests build is built here first, once.  GPU runs
    (`-m gpu`) stay in one process: they share one device.  MMB_TEST_WORKERS=0 turns this off, -n on the command line overrides it."""
    if hasattr(config, "workerinput") or os.environ.get("MMB_TEST_WORKERS") == "0":
        return None
    if (getattr(config.option, "markexpr", "") or "").replace(" ", "") != "notgpu":
        return None
    if getattr(config.option, "numprocesses", None) or not config.pluginmanager.hasplugin("xdist"):
        return None
    workers = int(os.environ.get("MMB_TEST_WORKERS", min(6, os.cpu_count() or 1)))
    if workers < 2 or getattr(config.option, "collectonly", False):
        return None
    try:
        _make("-j8", "lib", "simt")
        _make("-j8", "all")
        _make("port", cwd=os.path.join(ROOT, "oracle"))
        if os.path.isdir("/root/reference/src"):
            _make("ref", cwd=os.path.join(ROOT, "oracle"))
    except RuntimeError as e:                       # the fixtures report build problems test by test
        print(f"conftest: pre-build failed ({str(e)[:200]}); running in one process", file=sys.stderr)
        return None
    config.option.numprocesses = workers
    return None


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
