"""Parity tests proper: the CUDA library (sm_100a kernels) through the C ABI and the Python host layer,
against (1) the reference's recorded outputs, (2) the oracle on larger / richer synthetic cases,
(3) size-independent properties at benchmark-like sizes. fp64 tolerance 1e-10 relative (north star);
fp32 tolerance 1e-5 relative against the fp64 oracle (the reference's own fp32-vs-fp64
gap is ~1e-6)."""
import numpy as np
import pytest

from golden_util import Golden, available, relerr, relerr_cols, group_relerr, state_scales
from adfvm_b200 import function, cases
from oracle import adfvm_oracle as O

pytestmark = pytest.mark.gpu
CASES = [c for c in available() if not c.endswith("_fp32")]
TOL64, TOL32 = 1e-10, 1e-5


@pytest.mark.parametrize("name", CASES)
def test_golden_primal(name, cudalib):
    g = Golden(name)
    for run in ("orig", "perturb"):
        f = function.PrimalFunction(g.spec, np.float64)
        for ci, nm, inp, opt, out in g.calls(run, "primal"):
            r = f(*inp, **opt)
            for a, b in zip(r, out):
                assert relerr(a, b) < TOL64
                assert relerr_cols(a, b) < 10 * TOL64        # every component against its own size (1e-9)


@pytest.mark.parametrize("name", CASES)
def test_golden_adjoint(name, cudalib):
    g = Golden(name)
    f = function.PrimalFunction(g.spec, np.float64).grad()
    for ci, nm, inp, opt, out in g.calls("adjoint", "primal_grad"):
        r = f(*inp, **opt)
        sc = state_scales(inp)
        assert group_relerr(r[:3], out[:3], sc) < TOL64
        assert group_relerr(r[3:6], out[3:6], sc) < TOL64


def _adj_seed(case):
    rng = np.random.RandomState(3)
    return [np.ascontiguousarray(rng.randn(*s.shape) * w, case.dtype) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]


@pytest.mark.parametrize("make", [lambda dt: cases.periodic_box((16, 14, 12), dt, warp=0.03),
                                  lambda dt: cases.walled_box((12, 10, 6), dt)])
@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL64), (np.float32, TOL32)])
def test_oracle_parity(make, dtype, tol, cudalib):
    case = make(dtype)
    ref_case = make(np.float64)
    f = function.PrimalFunction(case.spec, dtype)
    state, ref_state = case.state, ref_case.state
    for step in range(3):
        out = f(*case.inputs(state), replace_reusable=(step == 0), return_reusable=True)
        ref = O.primal(ref_case.spec, ref_case.inputs(ref_state))
        for a, b in zip(out, ref):
            assert relerr(a, b) < tol, (step, relerr(a, b))
        state, ref_state = list(out[:3]), list(ref[:3])
    adj = _adj_seed(case)
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    gref = O.primal_grad(ref_case.spec, ref_case.adjoint_inputs(ref_case.state, [a.astype(np.float64) for a in adj]))
    sc = state_scales(ref_case.state)
    assert group_relerr(g[:3], gref[:3], sc) < tol
    assert group_relerr(g[3:6], gref[3:6], sc) < tol


# ---- the code path that is benchmarked (regular 4x4x8 tiles on lattice planes, the 444-tile L2 prefetch distance, several
# waves of CTAs) and the other two kernel variants, against the REFERENCE's own compiled step functions (oracle/_ref, built from
# the unmodified reference's generated code by oracle/ref_harness/build_ref_graph.py; the oracle port if it did not travel)
LARGE = [((64, 64, 64), 0.0, 0, 128),      # variant 0, (T,TS) = (128,288): regular tiles, 2048 tiles = 4.6 waves of 444 CTAs
         ((92, 92, 23), 0.0, 3, 64),       # ragged lattice (23 planes): 64-cell tiles (64,448)
         ((46, 46, 46), 0.0, 3, 64),
         ((40, 36, 28), 0.02, 2, 128),     # warped (no lattice): coordinate bisection, halo 193 -> (128,384)
         ((16, 14, 12), 0.03, 1, 128)]     # halo 190 -> (128,320)


def _reference_functions(case):
    from oracle import refgraph
    if refgraph.available("box_cyclic"):
        R = refgraph.IsolatedRefGraph("box_cyclic")       # one mesh per reference process (see oracle/refgraph.py)
        R.initialize(case.mesh)
        return R.primal, R.primal_grad, "reference"
    return (lambda *inp, **kw: O.primal(case.spec, list(inp))), (lambda *inp, **kw: O.primal_grad(case.spec, list(inp))), "port"


@pytest.mark.parametrize("n,warp,variant,tcells", LARGE)
def test_large_parity_against_reference(n, warp, variant, tcells, cudalib):
    case = cases.periodic_box(n, warp=warp)
    rp, rg, kind = _reference_functions(case)
    f = function.PrimalFunction(case.spec, np.float64)
    state = rstate = case.state
    for step in range(2):
        out = f(*case.inputs(state), replace_reusable=(step == 0), return_reusable=True)
        ref = rp(*case.inputs(rstate), replace_reusable=True)
        for a, b in zip(out, ref):
            assert relerr(a, b) < TOL64, (kind, step, relerr(a, b))
        state, rstate = list(out[:3]), [np.ascontiguousarray(x) for x in ref[:3]]
    assert f.tile_halo_stats()[1] == variant and f.tile_stats()[3] == tcells      # the kernel variant this case is meant to cover
    adj = _adj_seed(case)
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    gref = rg(*case.adjoint_inputs(case.state, adj))
    sc = state_scales(case.state)
    assert group_relerr(g[:3], gref[:3], sc) < TOL64
    assert group_relerr(g[3:6], gref[3:6], sc) < TOL64
    # component by component as well (a group norm hides the small ones): each array against its own max, 1e-9
    for a, b in zip(g, gref):
        assert relerr(a, b) < 1e-9


def test_large_parity_fp32(cudalib):
    """fp32 kernels at 64^3 against the fp64 reference, stated fp32 tolerance 1e-5"""
    case = cases.periodic_box(64, np.float32)
    case64 = cases.periodic_box(64)
    rp, rg, kind = _reference_functions(case64)
    f = function.PrimalFunction(case.spec, np.float32)
    out = f(*case.inputs(), replace_reusable=True)
    ref = rp(*case64.inputs(), replace_reusable=True)
    for a, b in zip(out, ref):
        assert relerr(a, b) < TOL32
    adj = _adj_seed(case)
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    gref = rg(*case64.adjoint_inputs(case64.state, [a.astype(np.float64) for a in adj]))
    sc = state_scales(case64.state)
    assert group_relerr(g[:3], gref[:3], sc) < TOL32 and group_relerr(g[3:6], gref[3:6], sc) < TOL32


def test_large_walled_parity(cudalib):
    """12 288-cell walled channel (every boundary condition class, 96 tiles) against the oracle"""
    case = cases.walled_box((32, 24, 16))
    f = function.PrimalFunction(case.spec, np.float64)
    out = f(*case.inputs(), replace_reusable=True)
    ref = O.primal(case.spec, case.inputs())
    for a, b in zip(out, ref):
        assert relerr(a, b) < TOL64
    adj = _adj_seed(case)
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    gref = O.primal_grad(case.spec, case.adjoint_inputs(case.state, adj))
    sc = state_scales(case.state)
    assert group_relerr(g[:3], gref[:3], sc) < TOL64 and group_relerr(g[3:6], gref[3:6], sc) < TOL64


def test_bitwise_deterministic(cudalib):
    """no float atomics anywhere: two runs give identical bits, primal and adjoint"""
    case = cases.periodic_box(24, warp=0.02)
    res = []
    for rep in range(2):
        f = function.PrimalFunction(case.spec, np.float64)
        out = f(*case.inputs(), replace_reusable=True)
        g = f.grad()(*case.adjoint_inputs(case.state, _adj_seed(case)))
        res.append([o.copy() for o in out] + [x.copy() for x in g])
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_graph_replay_matches_eager(cudalib):
    """whole steps run eagerly the first time, are captured into a CUDA graph the second time and replayed afterwards
    (per buffer rotation): same inputs must give identical bits in all three modes, primal and adjoint"""
    import torch
    case = cases.walled_box((10, 8, 4))
    stream = torch.cuda.Stream()                      # the legacy default stream cannot be captured
    f = function.PrimalFunction(case.spec, np.float64, stream=stream.cuda_stream)
    fa = f.grad()
    adj = _adj_seed(case)
    outs, grads = [], []
    for rep in range(8):
        out = f(*case.inputs(), replace_reusable=True, return_reusable=True)
        outs.append([o.copy() for o in out])
        g = fa(*case.adjoint_inputs(case.state, adj), return_static=True, zero_static=True)
        grads.append([x.copy() for x in g])
    for rep in range(1, 8):
        for a, b in zip(outs[0], outs[rep]):
            assert np.array_equal(a, b), rep
        for a, b in zip(grads[0], grads[rep]):
            assert np.array_equal(a, b), rep
    assert f.graph_replays >= 4, f.graph_replays          # the later repetitions really were replays
    # resident stepping (the replayed path) against one host round trip per step
    f2 = function.PrimalFunction(case.spec, np.float64)
    state = case.state
    for s in range(7):
        out = f2(*case.inputs(state), replace_reusable=True, return_reusable=True)
        state = list(out[:3])
    f3 = function.PrimalFunction(case.spec, np.float64, stream=stream.cuda_stream)
    f3(*case.inputs(), replace_reusable=True, return_reusable=False)
    for s in range(5):
        f3.step_resident(case.dt)
    out3 = f3(*case.inputs(), replace_reusable=False, return_reusable=True)
    assert f3.graph_replays >= 1
    for a, b in zip(out3, out):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", [c for c in available() if c.endswith("_fp32")])
def test_fp32_matches_reference_fp32_on_device(name, cudalib):
    """fp32 kernels against the reference's own fp32 mode (tests/test_fp32_golden.py), 2e-6 per call"""
    from test_fp32_golden import replay_fp32
    replay_fp32(name)


def test_overlapped_seed_upload_is_bitwise_equal(cudalib, monkeypatch):
    """primal_grad from host arrays uploads the adjoint seeds on a side stream while the forward sweep runs (large
    meshes); forced here on a small mesh and compared with the sequential path"""
    case = cases.walled_box((10, 8, 4))
    adj = _adj_seed(case)
    res = []
    for threshold in ("0", "1000000000"):
        monkeypatch.setenv("ADFVM_OVERLAP_MIN_CELLS", threshold)
        f = function.PrimalFunction(case.spec, np.float64)
        fa = f.grad()
        out = []
        for rep in range(3):
            g = fa(*case.adjoint_inputs(case.state, adj), return_static=True, zero_static=True)
            out.append([x.copy() for x in g])
        res.append(out)
    for a, b in zip(res[0], res[1]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_mesh_parameter_gradients_on_device(cudalib):
    """parameters = 'mesh': the ten metric-array gradients from the device against the oracle's reverse mode, 1e-10"""
    from oracle import adfvm_oracle as O
    for make in (lambda: cases.walled_box((6, 5, 3)), lambda: cases.cylinder2d(8, 10, dt=2e-9)):
        case = make()
        case.spec = dict(case.spec, parameters="mesh")
        f = function.PrimalFunction(case.spec, np.float64)
        adj = _adj_seed(case)
        g = f.grad()(*case.adjoint_inputs(case.state, adj), zero_static=True)
        gref = O.primal_grad(case.spec, case.adjoint_inputs(case.state, adj))
        assert len(g) == 13
        for a, b in zip(g[3:], gref[3:]):
            b = np.asarray(b).reshape(a.shape)
            assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()


def test_mesh_gradients_match_reference_on_device(cudalib):
    """fixture box_walls_mesh (the reference's own parameters='mesh' adjoint) replayed on the device, 1e-10 per array"""
    from test_mesh_param_golden import replay
    replay()


def test_mesh_metrics_on_device(cudalib):
    """adfvm_mesh_metrics on the device against the restatement of cmesh.cpp, fp64 1e-12"""
    from test_mesh_metrics_device import compare
    compare(cudalib)


def test_block_equals_host_round_trips_on_device(cudalib):
    """device-side checkpoint block (adfvm_primal_block / adfvm_adjoint_block) against one host round trip per step,
    bit for bit, with whole steps replayed as CUDA graphs"""
    import torch
    from test_block import block_vs_stepwise
    block_vs_stepwise(cudalib, stream=torch.cuda.Stream().cuda_stream)


def test_resident_equals_host_roundtrip(cudalib):
    case = cases.periodic_box(20)
    f1 = function.PrimalFunction(case.spec, np.float64)
    state = case.state
    for s in range(3):
        out = f1(*case.inputs(state), replace_reusable=True, return_reusable=True)
        state = list(out[:3])
    f2 = function.PrimalFunction(case.spec, np.float64)
    f2(*case.inputs(), replace_reusable=True, return_reusable=False)
    f2.step_resident(case.dt)
    out2 = f2(*case.inputs(), replace_reusable=False, return_reusable=True)
    for a, b in zip(out2, out):
        assert np.array_equal(a, b)


def test_adjoint_vs_finite_difference(cudalib):
    """dJ/dS from one adjoint step against a central finite difference of the primal step (objective on the
    stage-1 state + a linear functional of the output), the reference's end-to-end criterion
    (tests/test_adjoint.py:33) applied to a single step where FD noise is small."""
    case = cases.walled_box((10, 8, 4))
    f = function.PrimalFunction(case.spec, np.float64)
    adj = _adj_seed(case)
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    rng = np.random.RandomState(5)
    pert = [rng.randn(*s.shape) * w for s, w in zip(case.source, (1e1, 1e3, 1e6))]
    def J(sign, eps=1e-3):
        c2 = cases.walled_box((10, 8, 4))
        c2.source = [s + sign * eps * d for s, d in zip(c2.source, pert)]
        out = function.PrimalFunction(c2.spec, np.float64)(*c2.inputs(), replace_reusable=True)
        return sum(float((o * a).sum()) for o, a in zip(out[:3], adj)) + float(out[4][0, 0])
    fd = (J(+1) - J(-1)) / 2e-3
    ad = sum(float((a * d).sum()) for a, d in zip(g[3:6], pert))
    assert abs(fd - ad) / abs(fd) < 1e-6, (fd, ad)


def test_large_periodic_properties(cudalib):
    """64^3 periodic box: exact discrete conservation of mass/momentum/energy with zero source (every internal
    and cyclic face flux enters two cells with opposite sign), checked through the volume-weighted sums."""
    case = cases.periodic_box(64)
    case.source = [np.zeros_like(s) for s in case.source]
    f = function.PrimalFunction(case.spec, np.float64)
    out = f(*case.inputs(), replace_reusable=True)
    V = case.mesh.volumes
    for a, b in zip(out[:3], case.state):
        s0, s1 = (b * V).sum(axis=0), (a * V).sum(axis=0)
        assert np.all(np.abs(s1 - s0) <= 1e-12 * np.abs(b * V).sum(axis=0))
    assert np.isfinite(out[3]).all() and out[3][0, 0] > 0


# ---- adjoint artificial viscosity: `primal_grad_viscous` on the device
@pytest.mark.parametrize("name", ["box_walls", "box_cyclic", "cyl2d"])
@pytest.mark.parametrize("vt", ["abarbanel", "turkel", "uniform"])
def test_adjoint_viscosity_matches_reference_on_device(name, vt, cudalib):
    """M_2norm against the recording of the unmodified reference (oracle/ref_harness/gen_viscosity.py)"""
    import test_viscosity as tv
    g, z, inputs = tv._fixture(name)
    f = function.PrimalFunction(g.spec, np.float64)
    f(*inputs, replace_reusable=True)
    M = f.grad().viscous(vt).adjoint_viscosity(*inputs[:3], float(z["scaling"]))
    assert relerr(M, z["M_2norm_" + vt]) < TOL64


@pytest.mark.parametrize("name,vt,scaling", [("box_walls", "abarbanel", 3e4), ("box_cyclic", "turkel", 1e5), ("cyl2d", "abarbanel", 1e0)])
def test_viscous_adjoint_step_on_device(name, vt, scaling, cudalib):
    import test_viscosity as tv
    g, inp, r, out, expect, fa = tv.viscous_case(name, vt, scaling, None)
    sc = state_scales(inp)
    assert group_relerr(r[:3], expect, sc) < TOL64
    assert group_relerr(expect, out[:3], sc) > 1e-3
    assert group_relerr(r[3:6], out[3:6], sc) < TOL64


@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL64), (np.float32, 2e-4)])
def test_viscous_adjoint_large_against_oracle(dtype, tol, cudalib):
    """32^3 periodic box (256 tiles, several CTAs per SM): M_2norm and the smoothed adjoint against the oracle (numpy
    eigvalsh + direct sparse solve). fp32: the eigenvalue of a matrix with entries ~1e4 carries ~1e-4 relative error in
    single precision whatever the solver - the reference's own fp32 build (cusolver Ssyevj) is in the same position."""
    from oracle import adjoint_viscosity as AV
    case, ref_case = cases.periodic_box((32, 32, 32), dtype, warp=0.03), cases.periodic_box((32, 32, 32), np.float64, warp=0.03)
    scaling, vt = 3e2, "abarbanel"
    adj = _adj_seed(ref_case)
    inp = ref_case.adjoint_inputs(ref_case.state, adj); inp[-1] = np.array([[scaling]], np.float64)
    M, DT = AV.adjoint_viscosity(ref_case.spec, inp, vt, scaling)
    plain = O.primal_grad(ref_case.spec, ref_case.adjoint_inputs(ref_case.state, adj))
    expect = AV.apply_adjoint_viscosity(ref_case.spec, inp, DT, plain[:3])
    f = function.PrimalFunction(case.spec, dtype)
    f(*case.inputs(), replace_reusable=True)
    fa = f.grad().viscous(vt)
    dinp = case.adjoint_inputs(case.state, [a.astype(dtype) for a in adj]); dinp[-1] = np.array([[scaling]], dtype)
    r = fa(*dinp)
    Md = fa.adjoint_viscosity(*case.state, scaling)
    assert relerr(Md, M) < tol
    sc = state_scales(inp)
    assert group_relerr(expect, plain[:3], sc) > 1e-3
    assert group_relerr(r[:3], expect, sc) < tol
    assert 1 <= fa.viscosity_iterations <= 200


def test_state_cache_on_device(cudalib):
    """PrimalFunction(state_cache=n) on the device: cached device copies give bitwise the results of the upload path"""
    import test_block
    test_block.test_state_cache_opt_in(cudalib)


def test_viscous_blocks_on_device(cudalib):
    """adfvm_b200.blocks with the adjoint viscosity after every step (adfvm_adjoint_block_viscous) against the step-wise loop"""
    import test_viscosity as tv
    tv.test_viscous_blocks_equal_stepwise_driver_loop(cudalib)
