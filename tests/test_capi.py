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


def test_adjoint_getters_before_any_adjoint_step_fail_cleanly(hostsim):
    """adfvm_get_adjoint / fields() before any adjoint data exists must return an error code, not touch null buffers"""
    from adfvm_b200 import _lib, function, cases
    case = cases.periodic_box(6)
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    f(*case.inputs(), replace_reusable=True)
    with pytest.raises(_lib.AdfvmError):
        f.grad().fields()
    C_ = case.mesh.nInternalCells
    a, b, c = np.zeros((C_, 1)), np.zeros((C_, 3)), np.zeros((C_, 1))
    f.grad().set_fields(a, b, c)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    rc = hostsim.dll.adfvm_get_adjoint(f.c.ctx, p(a), p(b), p(c), p(a), None, None, 0)      # one gradient array without the others
    assert rc != 0


def test_primal_grad_leaves_the_resident_primal_state_alone(hostsim):
    """Function_primal's reuse buffers belong to `primal` (adpy/adpy/variable.py:382-388): a primal_grad call in between
    must not change what primal(replace_reusable=False) continues from"""
    from adfvm_b200 import function, cases
    case = cases.periodic_box((8, 6, 5), warp=0.02)
    adj = [np.ones_like(s) * w for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    f1 = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    f1(*case.inputs(), replace_reusable=True)
    ref = f1(*case.inputs(), replace_reusable=False)
    f2 = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    o = f2(*case.inputs(), replace_reusable=True)
    other = [s * 1.01 for s in case.state]
    f2.grad()(*case.adjoint_inputs(other, adj))
    out = f2(*case.inputs(), replace_reusable=False)
    for a, b in zip(out, ref):
        assert np.array_equal(a, b)
