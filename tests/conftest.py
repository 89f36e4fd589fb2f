import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle is compiled (gcc, seconds).  The CUDA library is built by
    __graft_entry__.build(); tests never build it implicitly on the GPU box."""
    so = ROOT / "oracle" / "libs2oracle.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    lib = ROOT / "synth2_b200" / "libs2cuda.so"
    if not lib.exists():
        from synth2_b200.build import build_lib
        build_lib()
    yield


def has_gpu():
    import ctypes
    from synth2_b200 import lib
    n = ctypes.c_int(0)
    rc = lib().s2_device_count(ctypes.byref(n))
    return rc == 0 and n.value > 0
