import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def build_hostsim():
    """CPU simulator of the device code (test infrastructure, see tests/hostsim/hostsim.cpp)."""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libadfvm_hostsim.so")
    csrc = os.path.join(ROOT, "adfvm_b200", "csrc")
    srcs = [os.path.join(d, "hostsim.cpp")] + [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".h", ".inc"))] + \
           [os.path.join(ROOT, "include", "adfvm_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", so, srcs[0]], cwd=d)
    return so


@pytest.fixture(scope="session")
def hostsim():
    from adfvm_b200 import _lib
    return _lib.Lib(build_hostsim())


@pytest.fixture(scope="session")
def cudalib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adfvm_b200 import _lib
    lib = _lib.default_lib()      # raises if the extension is missing: no fallback
    assert lib.is_cuda
    return lib
