"""Edge cases of the path through the C ABI (CPU simulator of the device code): meshes smaller than a warp, a single
cell, a 1-D row longer than a sub-tile, and a patch without faces - against the oracle, fp64 1e-10."""
import numpy as np
import pytest

from adfvm_b200 import cases, function, hexmesh
from adfvm_b200.metrics import build_mesh
from oracle import adfvm_oracle as O

TOL = 1e-10


def _check(case, lib):
    f = function.PrimalFunction(case.spec, np.float64, lib=lib)
    out = f(*case.inputs(), replace_reusable=True, return_reusable=True)
    ref = O.primal(case.spec, case.inputs())
    for a, b in zip(out, ref):
        assert np.abs(a - b).max() <= TOL * max(np.abs(b).max(), 1e-300)
    adj = [np.ones_like(s) * w for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    gref = O.primal_grad(case.spec, case.adjoint_inputs(case.state, adj))
    sc = [float(np.abs(s).max()) for s in case.state]
    for grp in (slice(0, 3), slice(3, 6)):
        num = max(np.abs(a - b).max() * s for a, b, s in zip(g[grp], gref[grp], sc))
        den = max(np.abs(b).max() * s for b, s in zip(gref[grp], sc))
        assert num <= TOL * den


@pytest.mark.parametrize("n", [(1, 1, 1), (2, 2, 2), (3, 1, 2), (33, 1, 1)])
def test_tiny_periodic_boxes(hostsim, n):
    _check(cases.periodic_box(n), hostsim)


def test_patch_without_faces(hostsim):
    """a patch of zero faces (OpenFOAM decompositions produce them) between the others"""
    lo, hi = (0., 0., 0.), (1., 1., 1.)
    poly = hexmesh.box_mesh((4, 3, 2), lo, hi, patches=[
        ("a_in", "patch", ["x-"], {}), ("b_none", "patch", [], {}), ("c_out", "patch", ["x+"], {}),
        ("walls", "symmetryPlane", ["y-", "y+"], {}),
        ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}), ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    mesh = build_mesh(poly)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    U, T, p = cases.smooth_state(cc)
    k0 = {"keys": []}
    zg = dict(type="zeroGradient", **k0)
    cyc = {"z1": dict(type="cyclic", **k0), "z2": dict(type="cyclic", **k0)}
    bcs = {f: dict(cyc, a_in=zg, b_none=zg, c_out=zg, walls=dict(type="symmetryPlane", **k0)) for f in ("U", "T", "p")}
    spec = cases._spec(mesh, bcs, {"kind": "patch_pA", "patch": "c_out"})
    case = cases.Case(mesh, spec, cases.conservative(U, T, p), cases.gaussian_source(cc), {}, 1e-6)
    assert mesh.boundary["b_none"]["nFaces"] == 0
    _check(case, hostsim)


def test_sum_T_objective(hostsim):
    """objective of reference templates/box.py:10-16 (sum of T over the cells, no volume weight)"""
    case = cases.periodic_box((6, 5, 4), warp=0.02)
    case.spec = dict(case.spec, objective={"kind": "cell_T"})
    _check(case, hostsim)


@pytest.mark.parametrize("par", [("BCs", "U", "lid", "value"), ("BCs", "T", "lid", "value"), ("BCs", "p", "outlet", "value"),
                                 ("BCs", "p", "inlet", "Tt"), ("BCs", "p", "inlet", "pt")])
def test_bc_parameter_gradients(hostsim, par):
    """adjoint with one BC input array as the parameter block (reference apps/adjoint.py:108-116; the reference-recorded
    fixture box_walls_bcpt pins the pt case, the others are checked against the oracle's autograd)"""
    case = cases.walled_box((6, 5, 3))
    case.spec = dict(case.spec, parameters=list(par))
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    adj = [np.ones_like(s) * w for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    gref = O.primal_grad(case.spec, case.adjoint_inputs(case.state, adj))
    assert len(g) == 4 and g[3].shape == gref[3].shape
    assert np.abs(gref[3]).max() > 0
    assert np.abs(g[3] - gref[3]).max() <= TOL * np.abs(gref[3]).max()
    # static accumulator: a second call without zero_static adds the same gradient again (adpy/adpy/variable.py:484-490)
    g2 = f.grad()(*case.adjoint_inputs(case.state, adj), zero_static=False)
    assert np.abs(g2[3] - 2 * gref[3]).max() <= 10 * TOL * np.abs(gref[3]).max()


@pytest.mark.parametrize("which", ["walls", "periodic", "cylinder", "step"])
def test_mesh_parameter_gradients(hostsim, which):
    """parameters = 'mesh' (reference apps/adjoint.py:105-107): gradient with respect to the ten metric arrays
    (areas, volumesL, volumesR, weights, deltas, normals, deltasUnit, linearWeights, quadraticWeights, volumes) against the
    oracle's reverse mode, every array to 1e-10 of its own size; BCs that read the normal (symmetryPlane, CBC_TOTAL_PT
    without direction), the drag / p.A / T.V objectives, Roe and Lax-Friedrichs boundary solvers, inviscid and viscous"""
    case = {"walls": lambda: cases.walled_box((6, 5, 3)), "periodic": lambda: cases.periodic_box((6, 5, 4), warp=0.03),
            "cylinder": lambda: cases.cylinder2d(8, 10, dt=2e-9), "step": lambda: cases.forward_step(20, 10, dt=1e-3)}[which]()
    case.spec = dict(case.spec, parameters="mesh")
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    adj = [np.ones_like(s) * w for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    g = f.grad()(*case.adjoint_inputs(case.state, adj), zero_static=True)
    gref = O.primal_grad(case.spec, case.adjoint_inputs(case.state, adj))
    assert len(g) == 13
    sc = [float(np.abs(s).max()) for s in case.state]
    num = max(np.abs(a - b).max() * s for a, b, s in zip(g[:3], gref[:3], sc))
    assert num <= TOL * max(np.abs(b).max() * s for b, s in zip(gref[:3], sc))
    for a, b in zip(g[3:], gref[3:]):
        b = np.asarray(b).reshape(a.shape)
        assert np.abs(a - b).max() <= TOL * np.abs(b).max()        # (an inviscid case has exactly zero deltas / deltasUnit gradients)
    assert sum(np.abs(np.asarray(b)).max() > 0 for b in gref[3:]) >= 8
    g2 = f.grad()(*case.adjoint_inputs(case.state, adj))          # accumulator was zeroed: same gradients again
    for a, b in zip(g2[3:], g[3:]):
        assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_holes(hostsim, seed):
    """irregular topologies: a 3-D box with random cells removed (ragged sub-tiles, cells with several boundary faces,
    possibly disconnected pieces) - the tile schedule must stay consistent (the simulator checks every entry) and the
    results must match the oracle"""
    rng = np.random.RandomState(seed)
    n = (7, 6, 5)
    keep = rng.rand(n[2], n[1], n[0]) > 0.25
    lo, hi = (0., 0., 0.), (1.4, 1.2, 1.0)
    poly = hexmesh.masked_box_mesh(n, lo, hi, keep, [
        ("inlet", "patch", ["x-"], {}), ("outlet", "patch", ["x+"], {}), ("walls", "symmetryPlane", ["y-", "y+"], {}),
        ("z1", "patch", ["z-"], {}), ("z2", "patch", ["z+"], {})], hole=("holes", "patch", {}))
    mesh = build_mesh(poly)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    U, T, p = cases.smooth_state(cc, lo, hi)
    k0 = {"keys": []}
    zg, sym = dict(type="zeroGradient", **k0), dict(type="symmetryPlane", **k0)
    bcs = {f: dict(inlet=zg, outlet=zg, walls=sym, z1=zg, z2=zg, holes=sym if f == "U" else zg) for f in ("U", "T", "p")}
    spec = cases._spec(mesh, bcs, {"kind": "patch_pA", "patch": "holes"})
    case = cases.Case(mesh, spec, cases.conservative(U, T, p), cases.gaussian_source(cc, (0.7, 0.6, 0.5)), {}, 1e-6)
    _check(case, hostsim)
