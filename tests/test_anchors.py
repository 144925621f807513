"""Reproducibility anchors of BASELINE.md section 5 at their own sizes (48^3 periodic box, 4 steps; 500-cell shock tube,
20 steps), recorded from the UNMODIFIED reference (oracle/ref_harness/gen_golden.py generate_anchor ->
tests/golden/anchor_*.{npz,json}: objective.txt, per-step time series, strided samples + norms of the final fields).

tests/minidriver.py restates the drivers' time loops; the same loops run the oracle, the CPU simulator of the device code
and - in the `gpu` tests - the CUDA library, and must reproduce the reference's objective.txt: `orig` to 1e-10, the
adjoint sensitivity to 1e-9 (the north-star tolerances), the finite difference `perturb` to 2e-12 of the objective it is a
difference of, and the final state / final adjoint fields at the sampled cells to 1e-10."""
import numpy as np
import pytest

import minidriver
from golden_util import Anchor
from adfvm_b200 import cases, function
from oracle import adfvm_oracle as O

NAMES = ("rho", "rhoU", "rhoE")
ANAMES = ("rhoa", "rhoUa", "rhoEa")


def build(name):
    a = Anchor(name)
    b = a.meta["builder"]
    case = {"periodic_box": lambda: cases.periodic_box(tuple(b["n"])), "tube": lambda: cases.shock_tube(b["n"], b["width"]),
            "forward_step_shipped": cases.forward_step_shipped, "cylinder_shipped": cases.cylinder_shipped,
            "vane_cascade": lambda: cases.vane_cascade(nz=b["nz"])}[b["kind"]]()
    # the static description the reference reported for its own objects must be the one our builder produces
    assert [(p["name"], p["type"], p["startFace"], p["nFaces"]) for p in a.spec["patches"]] == \
           [(p["name"], p["type"], p["startFace"], p["nFaces"]) for p in case.spec["patches"]]
    assert a.spec["mu"] == case.spec["mu"] and a.spec["BCs"] == case.spec["BCs"]
    case.spec = a.spec
    case.dt = a.cf["dt"]
    a.check_fields("initial", NAMES, case.state, 1e-13)
    return a, case


def replay(a, case, primal, grad, tol_obj=1e-10, tol_adj=1e-9, tol_field=1e-10):
    nSteps, wi = a.cf["nSteps"], a.cf["writeInterval"]
    pert = case.source                                   # the case builders carry the perturbation of the source terms
    orig, state, series = minidriver.run_primal(primal, case, nSteps)
    assert abs(orig - a.objective["orig"]) <= tol_obj * abs(a.objective["orig"])
    assert np.allclose(series, a.z["timeSeries_orig"], rtol=tol_obj, atol=0)
    a.check_fields("orig", NAMES, state, tol_field)
    pval, pstate, _ = minidriver.run_primal(primal, case, nSteps, source=pert)
    assert abs((pval - orig) - a.objective["perturb"]) <= 2e-12 * abs(a.objective["orig"])
    a.check_fields("perturb", NAMES, pstate, tol_field)
    adj, fields, sens = minidriver.run_adjoint(primal, grad, case, nSteps, wi, pert)
    assert abs(adj - a.objective["adjoint"]) <= tol_adj * abs(a.objective["adjoint"]), (adj, a.objective["adjoint"])
    assert np.allclose(sens, a.z["sensTimeSeries"], rtol=tol_adj, atol=tol_adj * np.abs(a.z["sensTimeSeries"]).max())
    sc = [float(np.abs(s).max()) for s in case.state]
    a.check_fields("adjoint", ANAMES, fields, tol_field, sc)
    # the reference's own end-to-end criterion: adjoint sensitivity vs finite difference, 1e-3 (tests/test_adjoint.py:33)
    assert abs(adj - (pval - orig)) < 1e-3 * abs(adj)


def oracle_functions(case):
    return (lambda *inp, **kw: O.primal(case.spec, list(inp))), (lambda *inp, **kw: O.primal_grad(case.spec, list(inp)))


def test_anchor_tube500_oracle():
    a, case = build("anchor_tube500")
    replay(a, case, *oracle_functions(case))


def test_anchor_tube500_hostsim(hostsim):
    a, case = build("anchor_tube500")
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    replay(a, case, f, f.grad())


def test_anchor_box48_hostsim(hostsim):
    a, case = build("anchor_box48")
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    replay(a, case, f, f.grad())


# BASELINE.json configs 2 and 3 on the meshes the reference ships (cases/*/constant/polyMesh/blockMeshDict through
# adfvm_b200.blockmesh), recorded from the unmodified reference at full size
@pytest.mark.parametrize("name", ["anchor_forwardstep", "anchor_cylinder", "anchor_vane"])
def test_anchor_shipped_mesh_hostsim(name, hostsim):
    a, case = build(name)
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    replay(a, case, f, f.grad())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["anchor_tube500", "anchor_box48", "anchor_forwardstep", "anchor_cylinder", "anchor_vane"])
def test_anchor_on_device(name, cudalib):
    a, case = build(name)
    f = function.PrimalFunction(case.spec, np.float64)
    replay(a, case, f, f.grad())


# ---- the device-side checkpoint block against the REFERENCE's recorded run (not against the step-wise
# path of the same library): Solver.run(mode='forward') + the backward loop of Adjoint.run with every state, the adjoint fields
# and the source-gradient accumulator resident; the host sees one state per checkpoint and the accumulated gradient per block
def replay_blocks(a, case, f, tol_obj=1e-10, tol_adj=1e-9, tol_field=1e-10):
    from adfvm_b200 import blocks
    nSteps, wi = a.cf["nSteps"], a.cf["writeInterval"]
    C = case.mesh.nInternalCells
    pert = case.source                                   # (the context below zeroes the case's source terms)
    with minidriver.source_terms(case, None):
        series, checkpoints = blocks.forward_blocks(f, case.inputs(), nSteps, wi, case.dt)
        orig = sum(series) / nSteps
        assert abs(orig - a.objective["orig"]) <= tol_obj * abs(a.objective["orig"])
        assert np.allclose(series, a.z["timeSeries_orig"], rtol=tol_obj, atol=0)
        a.check_fields("orig", NAMES, checkpoints[-1], tol_field)
        zero = [np.zeros((C, 1)), np.zeros((C, 3)), np.zeros((C, 1))]
        adj, fields = blocks.adjoint_blocks(f, f.grad(), case.inputs, checkpoints, nSteps, wi, case.dt, zero, pert)
        assert abs(adj - a.objective["adjoint"]) <= tol_adj * abs(a.objective["adjoint"]), (adj, a.objective["adjoint"])
        a.check_fields("adjoint", ANAMES, fields, tol_field, [float(np.abs(s).max()) for s in case.state])


@pytest.mark.parametrize("name", ["anchor_tube500", "anchor_box48"])
def test_anchor_blocks_hostsim(name, hostsim):
    a, case = build(name)
    replay_blocks(a, case, function.PrimalFunction(case.spec, np.float64, lib=hostsim))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["anchor_tube500", "anchor_box48", "anchor_cylinder"])
def test_anchor_blocks_on_device(name, cudalib):
    a, case = build(name)
    replay_blocks(a, case, function.PrimalFunction(case.spec, np.float64))
