"""The C-ABI product library: builds for sm_100a, loads without a GPU, exports every symbol the header
declares, and fails loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "adfvm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adfvm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_header_symbols():
    import __graft_entry__ as ge
    ge.build()
    from adfvm_b200 import _lib
    lib = _lib.default_lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib.dll, s), s
    assert sorted(_lib.EXPORTS) == syms
    assert lib.is_cuda
    assert lib.dll.adfvm_version() == 100


def test_patch_struct_matches_header():
    from adfvm_b200 import _lib
    assert ctypes.sizeof(_lib.Patch) == 10 * 4


def test_no_cpu_fallback():
    """Without a CUDA device context creation must raise, not silently compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from adfvm_b200 import _lib, function, cases
    case = cases.periodic_box(4)
    with pytest.raises(_lib.AdfvmError):
        function.PrimalFunction(case.spec, np.float64, device=0)


def test_sass_is_sm100a():
    import subprocess
    from adfvm_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
