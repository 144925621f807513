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
