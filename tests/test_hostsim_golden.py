"""The DEVICE code (adfvm_b200/csrc/fvm_*.h bodies, step orchestration, C ABI) compiled for the CPU by
tests/hostsim, driven through the same Python host layer as the product, against the reference's recorded
outputs. This is the GPU-less check of the arithmetic, indexing and reverse sweep; the same comparisons run
through the CUDA library in test_gpu_parity.py (-m gpu). fp64 tolerance 1e-10 (north star)."""
import numpy as np
import pytest

from golden_util import Golden, available, relerr, relerr_cols, group_relerr, state_scales
from adfvm_b200 import function

CASES = [c for c in available() if not c.endswith("_fp32")]
TOL = 1e-10


@pytest.mark.parametrize("name", CASES)
def test_primal_calls(name, hostsim):
    g = Golden(name)
    for run in ("orig", "perturb"):
        f = function.PrimalFunction(g.spec, np.float64, lib=hostsim)
        for ci, nm, inp, opt, out in g.calls(run, "primal"):
            r = f(*inp, **opt)
            for a, b in zip(r, out):
                assert relerr(a, b) < TOL
                assert relerr_cols(a, b) < 10 * TOL          # every component against its own size (1e-9)


@pytest.mark.parametrize("name", CASES)
def test_adjoint_calls(name, hostsim):
    g = Golden(name)
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim).grad()
    for ci, nm, inp, opt, out in g.calls("adjoint", "primal_grad"):
        r = f(*inp, **opt)
        sc = state_scales(inp)
        assert group_relerr(r[:3], out[:3], sc) < TOL
        assert group_relerr(r[3:6], out[3:6], sc) < TOL


def test_reusable_state_semantics(hostsim):
    """state stays resident between calls; arrays passed without replace_reusable are ignored
    (adpy/adpy/variable.py:382-388); outputs are None unless return_reusable (:491-496)."""
    g = Golden("box_cyclic")
    calls = list(g.calls("orig", "primal"))
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim)
    _, _, inp0, _, out0 = calls[0]
    r0 = f(*inp0, replace_reusable=True, return_reusable=False)
    assert r0[0] is None and r0[1] is None and r0[2] is None
    _, _, inp1, _, out1 = calls[1]
    junk = [np.full_like(a, np.nan) for a in inp1[:3]] + list(inp1[3:])
    r1 = f(*junk, replace_reusable=False, return_reusable=True)
    for a, b in zip(r1, out1):
        assert relerr(a, b) < TOL


def test_static_gradient_accumulates(hostsim):
    """source-term gradients are static accumulators: summed across calls until zero_static
    (apps/adjoint.py:281-284, adpy/adpy/variable.py:484-490)."""
    g = Golden("tube")
    calls = list(g.calls("adjoint", "primal_grad"))[:2]
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim).grad()
    r0 = f(*calls[0][2], return_static=False, zero_static=False)
    assert r0[3] is None
    r1 = f(*calls[1][2], return_static=True, zero_static=True)
    expect = [a + b for a, b in zip(calls[0][4][3:6], calls[1][4][3:6])]
    assert group_relerr(r1[3:6], expect, state_scales(calls[0][2])) < TOL
    r2 = f(*calls[1][2], return_static=True, zero_static=True)
    assert group_relerr(r2[3:6], calls[1][4][3:6], state_scales(calls[0][2])) < TOL


def test_argument_validation(hostsim):
    g = Golden("tube")
    _, _, inp, opt, _ = next(g.calls("orig", "primal"))
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim)
    bad = list(inp); bad[1] = bad[1].astype(np.float32)
    with pytest.raises(TypeError):
        f(*bad, **opt)
    bad = list(inp); bad[9] = np.asfortranarray(np.tile(bad[9], (1, 1)))[:, ::-1]
    with pytest.raises(ValueError):        # static arrays are validated when they are uploaded (first call)
        function.PrimalFunction(g.spec, np.float64, lib=hostsim)(*bad, **opt)
    with pytest.raises(TypeError):
        f(*inp, bogus_option=True)
    bad = list(inp); bad[5] = bad[5] * 1.5          # volumesL inconsistent with volumes
    with pytest.raises(Exception):
        function.PrimalFunction(g.spec, np.float64, lib=hostsim)(*bad, **opt)
