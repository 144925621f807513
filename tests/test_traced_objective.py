"""The objective callback of the library (include/adfvm_b200.h adfvm_set_objective_callback) + the trace interpreter
(adfvm_b200/adpy_objective.py) against the library's native objectives on the same cases: a hand-built trace of
sum p_ghost*area over a patch (= native patch_pA) and of sum T*V over the cells (= native cell_TV) must give the same objective,
adjoint fields and source-term gradients. On the CPU simulator here, on the device in the `gpu` test (where the reference's
front-end, which produces real traces - tests/test_dropin_overlay.py - does not exist)."""
import numpy as np
import pytest

import fake_trace as ft
from adfvm_b200 import cases, function, adpy_objective
from golden_util import relerr, group_relerr, state_scales


def _traced_spec(case, kind):
    inp = case.inputs()
    ivars = [ft.Variable((1,)) for _ in inp]                  # one Variable per positional input of `primal`
    fields = [ft.Variable((case.mesh.nCells, 3)), ft.Variable((case.mesh.nCells, 1)), ft.Variable((case.mesh.nCells, 1))]
    if kind == "patch_pA":
        p = case.mesh.boundary[case.spec["objective"]["patch"]]
        obj = ft.patch_pressure_force(ivars, fields, ivars[4], ivars[4 + 11], p["startFace"], p["nFaces"])   # areas, neighbour
    else:
        obj = ft.cell_TV(fields, ivars[4 + 9], case.mesh.nInternalCells)                                       # volumes
    traced = adpy_objective.TracedObjective(ft.Function(ivars), fields, obj)
    return dict(case.spec, objective={"kind": "traced", "traced": traced})


def _compare(case, kind, lib=None):
    adj = [np.ones_like(s) * w for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    res = []
    for spec in (case.spec, _traced_spec(case, kind)):
        f = function.PrimalFunction(spec, np.float64, lib=lib)
        out = f(*case.inputs(), replace_reusable=True)
        g = f.grad()(*case.adjoint_inputs(case.state, adj, obja=0.7))
        res.append((out, g))
    (o1, g1), (o2, g2) = res
    assert abs(o1[4][0, 0]) > 0
    for a, b in zip(o2, o1):
        assert relerr(a, b) < 1e-13
    sc = state_scales(case.state)
    assert group_relerr(g2[:3], g1[:3], sc) < 1e-12 and group_relerr(g2[3:6], g1[3:6], sc) < 1e-12


def test_traced_patch_objective_matches_native(hostsim):
    _compare(cases.forward_step(30, 10), "patch_pA", hostsim)


def test_traced_cell_objective_matches_native(hostsim):
    _compare(cases.periodic_box((8, 6, 5), warp=0.02), "cell_TV", hostsim)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["patch_pA", "cell_TV"])
def test_traced_objective_on_device(kind, cudalib):
    _compare(cases.forward_step(60, 20) if kind == "patch_pA" else cases.periodic_box(16, warp=0.02), kind)
