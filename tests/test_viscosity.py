"""Adjoint artificial viscosity (`primal_grad_viscous`, apps/adjoint.py:127-141).

  1. the oracle's M_2norm / DT against recordings of the UNMODIFIED reference (tests/golden/visc_*.npz, made by
     oracle/ref_harness/gen_viscosity.py: the reference's own traced kernels + LAPACK dsyev): three cases x three types;
  2. the device code (CPU simulator here, CUDA library in test_gpu_parity.py) against those recordings (M_2norm) and against
     the oracle's direct sparse solve of the system matop_petsc.cpp assembles (the smoothed adjoint fields);
  3. decomposition invariance on thread ranks (the reference's criterion for multi-rank runs, tests/test_parallel.py:63-81).
Tolerance 1e-10 relative (fp64 north star); the fixtures are met to ~1e-14."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from golden_util import GOLDEN, Golden, group_relerr, relerr, state_scales
from adfvm_b200 import decompose, function
from oracle import adjoint_viscosity as AV

CASES = ["box_walls", "box_cyclic", "cyl2d"]
TOL = 1e-10


def _fixture(name):
    g = Golden(name)
    z = np.load(os.path.join(GOLDEN, "visc_%s.npz" % name))
    inputs = next(g.calls("orig", "primal"))[2]
    assert np.array_equal(inputs[0], z["rho"]) and np.array_equal(inputs[1], z["rhoU"]) and np.array_equal(inputs[2], z["rhoE"])
    return g, z, inputs


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("vt", AV.TYPES)
def test_oracle_matches_reference(name, vt):
    g, z, inputs = _fixture(name)
    M, DT = AV.adjoint_viscosity(g.spec, inputs, vt, float(z["scaling"]))
    assert relerr(M, z["M_2norm_" + vt]) < 1e-12
    assert relerr(DT, z["DT_" + vt]) < 1e-12


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("vt", AV.TYPES)
def test_device_code_matches_reference(name, vt, hostsim):
    g, z, inputs = _fixture(name)
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim)
    f(*inputs, replace_reusable=True)                    # uploads the static data
    M = f.grad().viscous(vt).adjoint_viscosity(*inputs[:3], float(z["scaling"]))
    assert relerr(M, z["M_2norm_" + vt]) < TOL


def viscous_case(name, vt, scaling, lib, dtype=np.float64, call=0):
    """one recorded primal_grad call of the reference replayed with the viscous function; returns (result, oracle result)"""
    g = Golden(name)
    calls = list(g.calls("adjoint", "primal_grad"))
    _, _, inp, opt, out = calls[call]
    inp = list(inp)
    inp[-1] = np.array([[scaling]], np.float64)
    M, DT = AV.adjoint_viscosity(g.spec, inp, vt, scaling)
    expect = AV.apply_adjoint_viscosity(g.spec, inp, DT, out[:3])
    if dtype != np.float64:
        inp = [a.astype(dtype) if isinstance(a, np.ndarray) and a.dtype == np.float64 else a for a in inp]
    fa = (function.PrimalFunction(g.spec, dtype, lib=lib) if lib is not None else function.PrimalFunction(g.spec, dtype)).grad().viscous(vt)
    r = fa(*inp, **opt)
    viscous_case.opt = opt
    return g, inp, r, out, expect, fa


@pytest.mark.parametrize("name,vt,scaling", [("box_walls", "abarbanel", 3e4), ("box_cyclic", "turkel", 1e5), ("cyl2d", "abarbanel", 1e0),
                                             ("box_walls", "uniform", 1e5), ("tube", "turkel", 1e2)])
def test_viscous_adjoint_step(name, vt, scaling, hostsim):
    g, inp, r, out, expect, fa = viscous_case(name, vt, scaling, hostsim)
    sc = state_scales(inp)
    assert group_relerr(r[:3], expect, sc) < TOL
    assert group_relerr(r[3:6], out[3:6], sc) < TOL                      # the parameter gradient is untouched by the smoothing
    # the smoothing really acted (otherwise the comparison above would only repeat test_adjoint_calls)
    assert group_relerr(expect, out[:3], sc) > 1e-3
    assert 1 <= fa.viscosity_iterations <= 500
    # diffusion conserves the volume integral of each field: the sums of the volume-weighted adjoint fields equal those of the
    # unsmoothed result of the same library (its own rounding noise included, which is all the rho adjoint of `tube` consists of)
    plain = fa.primal.grad()(*inp, **viscous_case.opt)
    for a, b in zip(r[:3], plain[:3]):
        assert np.allclose(a.sum(axis=0), b.sum(axis=0), rtol=1e-9, atol=1e-12 * np.abs(b).sum())


def test_viscous_resident_matches_host_call(hostsim):
    """viscous_resident after step_resident == the viscous call with host arrays"""
    g, inp, r, out, expect, fa = viscous_case("box_walls", "abarbanel", 3e4, hostsim)
    f = fa.primal
    f.set_state(*inp[:-6])
    plain = f.grad()
    rest = inp[-6:]
    plain.set_fields(*rest[:3])
    # set_fields loads A[0]; chain=True makes it this step's input
    plain.step_resident(float(inp[3][0, 0]), float(rest[4][0, 0]), chain=True)
    fa.viscous_resident(float(inp[3][0, 0]), 3e4)
    got = plain.fields(return_static=False)[:3]
    assert group_relerr(got, r[:3], state_scales(inp)) < 1e-12


def test_unknown_type_raises(hostsim):
    g = Golden("tube")
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim)
    with pytest.raises(NotImplementedError):
        f.grad().viscous("entropy_hughes")


@pytest.mark.parametrize("world", [2, 4])
def test_threads_decomposition_invariance(world, hostsim):
    """M_2norm normalisation (two all-reduced sums), its halo, and the diffusion solve across processor patches: every rank's
    smoothed adjoint equals the single-rank run on the undecomposed mesh. A WALLED box: the reference couples cells across
    processor / processorCyclic patches (matop_petsc.cpp:23-34,386-399) but not across cyclic ones (cellNeighboursMatOp = -1
    for every ghost cell, cmesh.cpp:244-267), so on periodic meshes its own result depends on the decomposition."""
    from adfvm_b200 import cases
    vt, scaling = "abarbanel", 3e4
    gl = cases.walled_box((8, 6, 4), warp=0.0)
    rng = np.random.RandomState(5)
    adj = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(gl.state, (1.0, 1e-2, 1e-5))]

    def scaled(inputs):
        inputs = list(inputs); inputs[-1] = np.array([[scaling]], np.float64); return inputs
    f = function.PrimalFunction(gl.spec, np.float64, lib=hostsim)
    f(*gl.inputs(), replace_reusable=True)
    fa = f.grad().viscous(vt)
    ref = fa(*scaled(gl.adjoint_inputs(gl.state, adj)))
    refM = fa.adjoint_viscosity(*gl.state, scaling)
    plain = f.grad()(*gl.adjoint_inputs(gl.state, adj))
    sc = [float(np.abs(s).max()) for s in gl.state]
    assert group_relerr(ref[:3], plain[:3], sc) > 1e-3
    parts = decompose.rank_cases(gl, world)
    uid = C.create_string_buffer(128)
    hostsim.check(hostsim.dll.adfvm_comm_unique_id(uid))
    results, errors = {}, []

    def run(rank):
        try:
            case, ids = parts[rank]
            fr = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
            fr.c.attach_comm(uid.raw, rank, world)
            fr(*case.inputs(), replace_reusable=True)
            a = [np.ascontiguousarray(x[ids]) for x in adj]
            far = fr.grad().viscous(vt)
            out = far(*scaled(case.adjoint_inputs(case.state, a)))
            M = far.adjoint_viscosity(*case.state, scaling)
            results[rank] = (ids, out, M)
        except Exception as e:      # pragma: no cover
            errors.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errors, errors
    for rank in range(world):
        ids, out, M = results[rank]
        assert relerr(M[:len(ids)], refM[ids]) < TOL
        num = max(np.abs(a - b[ids]).max() * s for a, b, s in zip(out[:3], ref[:3], sc))
        den = max(np.abs(b).max() * s for b, s in zip(ref[:3], sc))
        assert num / den < TOL


# ---- the drivers' time loop (tests/minidriver.py = Adjoint.run with viscousInterval = 1): oracle against the device code
def oracle_viscous(case, vt):
    from oracle import adfvm_oracle as O

    def f(*inp, **kw):
        inp = list(inp)
        out = O.primal_grad(case.spec, inp)
        M, DT = AV.adjoint_viscosity(case.spec, inp, vt, float(inp[-1][0, 0]))
        return AV.apply_adjoint_viscosity(case.spec, inp, DT, out[:3]) + list(out[3:])
    return f


def test_adjoint_run_with_viscosity_oracle_vs_device_code(hostsim):
    import minidriver
    from adfvm_b200 import cases
    from oracle import adfvm_oracle as O
    case = cases.walled_box((6, 5, 4))
    vt, scaling, nSteps, wi = "abarbanel", 3e4, 4, 2
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    got, gfields, gsens = minidriver.run_adjoint(f, f.grad().viscous(vt), case, nSteps, wi, case.source, scaling=scaling)
    oprimal = lambda *inp, **kw: O.primal(case.spec, list(inp))
    ref, rfields, rsens = minidriver.run_adjoint(oprimal, oracle_viscous(case, vt), case, nSteps, wi, case.source, scaling=scaling)
    plain, _, _ = minidriver.run_adjoint(f, f.grad(), case, nSteps, wi, case.source)
    assert abs(ref - plain) > 1e-3 * abs(plain)                       # the smoothing changes the sensitivity
    assert abs(got - ref) <= 1e-9 * abs(ref)
    assert np.allclose(gsens, rsens, rtol=1e-9, atol=1e-9 * np.abs(rsens).max())
    assert group_relerr(gfields, rfields, [float(np.abs(s).max()) for s in case.state]) < TOL


def test_viscous_blocks_equal_stepwise_driver_loop(hostsim):
    """adfvm_b200.blocks with the viscous function (states, adjoint fields and gradient accumulator resident, the smoothing after
    every step) against the step-wise loop with host round trips"""
    import minidriver
    from adfvm_b200 import blocks, cases
    case = cases.walled_box((6, 5, 4))
    vt, scaling, nSteps, wi = "turkel", 1e5, 4, 2
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    ref, rfields, _ = minidriver.run_adjoint(f, f.grad().viscous(vt), case, nSteps, wi, case.source, scaling=scaling)
    C_ = case.mesh.nInternalCells
    pert = case.source
    with minidriver.source_terms(case, None):
        f2 = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
        series, checkpoints = blocks.forward_blocks(f2, case.inputs(), nSteps, wi, case.dt)
        zero = [np.zeros((C_, 1)), np.zeros((C_, 3)), np.zeros((C_, 1))]
        fa2 = f2.grad()
        got, gfields = blocks.adjoint_blocks(f2, fa2, case.inputs, checkpoints, nSteps, wi, case.dt, zero, pert,
                                             viscous=fa2.viscous(vt), scaling=scaling)
    assert abs(got - ref) <= 1e-12 * abs(ref)
    assert group_relerr(gfields, rfields, [float(np.abs(s).max()) for s in case.state]) < 1e-12


# ---- the unmodified apps/adjoint.py with a case file that sets adjParams, served through the adpy overlay
REF = os.environ.get("ADFVM_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")
def test_reference_adjoint_driver_with_viscosity_through_overlay(hostsim):
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle", "ref_harness"))
    import gen_golden
    vt, scaling = "abarbanel", 3e4
    base = gen_golden.CASES["box_walls"]
    gen_golden.CASES["box_walls_visc"] = lambda: dict(base(), param_block="adjParams = [%r, %r, None]\nviscousInterval = 1\n" % (scaling, vt))
    c, case, casefile = gen_golden.write_case("box_walls_visc", "box_walls_visc_overlay")
    env = dict(os.environ, ADFVM_DROPIN_LIB=hostsim.path)
    runner = os.path.join(root, "oracle", "ref_harness", "run_overlay.py")
    for app in ("problem", "adjoint"):
        out = subprocess.run([sys.executable, runner, app, "--", casefile], cwd=case, env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    served = [ln for ln in out.stdout.split("\n") if ln.startswith("[overlay]")][-1]
    assert int(served.split()[-1]) == c["nSteps"], served          # every adjoint step went through primal_grad_viscous
    got = {ln.split()[0]: float(ln.split()[3]) for ln in open(os.path.join(case, "objective.txt")).read().strip().split("\n")}
    # the same steps replayed from the recording of the stock reference run (states, options), adjoint fields chained
    g = Golden("box_walls")
    cf = g.meta["casefile"]
    cc = g.mesh_array("adjoint", "cellCentres")
    C_ = len(next(g.calls("adjoint", "primal_grad"))[2][0])
    G = float(cf["amp"]) * np.exp(-float(cf["width"]) * np.linalg.norm(cc[:C_] - np.array(json.loads(cf["mid"])), axis=1, keepdims=True) ** 2)
    pert = [G, np.concatenate([G * 100, 0 * G, 0 * G], axis=1), G * 2e5]
    fa = function.PrimalFunction(g.spec, np.float64, lib=hostsim).grad().viscous(vt)
    adj, total = None, 0.0
    for ci, nm, inp, opt, o in g.calls("adjoint", "primal_grad"):
        inp = list(inp)
        if adj is not None:
            inp[-6:-3] = adj
        inp[-1] = np.array([[scaling]])
        r = fa(*inp, **opt)
        adj = [np.array(a, copy=True) for a in r[:3]]
        total += sum(float((a * b).sum()) for a, b in zip(r[3:6], pert))
    expect = total / c["nSteps"]
    ref_plain = {ln.split()[0]: float(ln.split()[3]) for ln in g.meta["objective_txt"]}["adjoint"]
    assert abs(expect - ref_plain) > 1e-3 * abs(ref_plain)
    assert abs(got["adjoint"] - expect) <= 1e-10 * abs(expect), (got, expect)
