"""fp32 against the reference's OWN fp32 mode: the fixtures box_cyclic_fp32 / box_walls_fp32 were recorded from the
unmodified reference built with scalar = float (oracle/ref_harness/gen_golden.py --fp32: cmesh_gpu natives,
config.precision = float32, SMALL = 1e-9 - the reference's `-g` default without --gpu_double, fp32 state AND fp32 mesh
metrics). Every recorded `primal` / `primal_grad` call is replayed in fp32 through the C ABI (CPU simulator here, the
device in tests/test_gpu_parity.py) and must agree to 2e-6 relative per call (measured 2e-7; both sides carry fp32
round-off of different summation orders) - the single-step bound of SURVEY section 8(c). The fp64 oracle comparison at
1e-5 lives in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from golden_util import Golden, available, group_relerr, relerr, state_scales
from adfvm_b200 import function

CASES32 = [c for c in available() if c.endswith("_fp32")]
TOL = 2e-6


def replay_fp32(name, **kw):
    g = Golden(name)
    n = 0
    for run in ("orig", "perturb"):
        f = function.PrimalFunction(g.spec, np.float32, **kw)       # static inputs (source terms) differ between the runs
        for ci, nm, inp, opt, out in g.calls(run, "primal"):
            r = f(*inp, **opt)
            if out[0] is not None:
                assert group_relerr(r[:3], out[:3], state_scales(inp)) < TOL
            assert relerr(r[3], out[3]) < TOL
            # objectives that are sums with cancellation (drag) lose digits in fp32 on both sides
            assert relerr(r[4], out[4]) < 50 * TOL
            n += 1
    fa = function.PrimalFunction(g.spec, np.float32, **kw).grad()
    for ci, nm, inp, opt, out in g.calls("adjoint", "primal_grad"):
        r = fa(*inp, **opt)
        sc = state_scales(inp)
        assert group_relerr(r[:3], out[:3], sc) < TOL
        assert group_relerr(r[3:6], out[3:6], sc) < TOL
        n += 1
    assert n >= 8


@pytest.mark.parametrize("name", CASES32)
def test_fp32_matches_reference_fp32(name, hostsim):
    assert CASES32
    replay_fp32(name, lib=hostsim)


@pytest.mark.parametrize("make", [lambda: __import__("adfvm_b200.cases", fromlist=["x"]).periodic_box((10, 8, 6), np.float32, warp=0.03),
                                  lambda: __import__("adfvm_b200.cases", fromlist=["x"]).walled_box((8, 6, 4), np.float32)])
def test_fp32_against_fp64_oracle(hostsim, make):
    """the stated fp32 tolerance of the north star: 1e-5 relative against the fp64 oracle (one step, primal and adjoint)"""
    from oracle import adfvm_oracle as O
    case = make()
    f = function.PrimalFunction(case.spec, np.float32, lib=hostsim)
    out = f(*case.inputs(), replace_reusable=True, return_reusable=True)
    inp64 = [np.asarray(a, np.float64) if isinstance(a, np.ndarray) and a.dtype == np.float32 else a for a in case.inputs()]
    ref = O.primal(case.spec, inp64)
    sc = state_scales(inp64)
    assert group_relerr(out[:3], ref[:3], sc) < 1e-5
    adj = [np.ascontiguousarray(np.ones_like(s) * w, np.float32) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    a64 = [np.asarray(a, np.float64) if isinstance(a, np.ndarray) and a.dtype == np.float32 else a for a in case.adjoint_inputs(case.state, adj)]
    gref = O.primal_grad(case.spec, a64)
    assert group_relerr(g[:3], gref[:3], sc) < 1e-5
    assert group_relerr(g[3:6], gref[3:6], sc) < 1e-5
