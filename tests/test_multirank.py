"""Multi-rank path on CPU: the rank-local step logic (processor patches, halo pack/unpack order, tags, reverse halo
with add, objective all-reduce) of the device code compiled for the CPU (tests/hostsim), checked for
decomposition invariance against the single-rank run on the undecomposed mesh — the reference's own criterion
(tests/test_parallel.py:32-41, 63-81: rel < 1e-9 after 10 steps; here 1e-10 after 2 primal steps + 1 adjoint step).
  * ranks as threads of one process (in-process rendezvous), 2 / 4 / 8 ranks;
  * ranks as 2 real processes over torch.distributed `gloo` (halo through host callbacks)."""
import ctypes as C
import os
import sys
import threading

import numpy as np
import pytest

from adfvm_b200 import decompose, function
from golden_util import relerr

TOL = 1e-10
N = (6, 5, 4)


def _seed(state):
    rng = np.random.RandomState(7)
    return [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(state, (1.0, 1e-2, 1e-5))]


def _single(world, lib, N=N):
    g = decompose.global_box(N, world)
    f = function.PrimalFunction(g.spec, np.float64, lib=lib)
    out = f(*g.inputs(), replace_reusable=True)
    out2 = f(*g.inputs(list(out[:3])), replace_reusable=True)
    adj = _seed(g.state)
    grad = f.grad()(*g.adjoint_inputs(g.state, adj))
    return g, out, out2, adj, grad


def _rank_run(rank, world, lib, uid, adj_global, results, attach, N=N, tiles=None):
    case = decompose.periodic_box_rank(N, rank, world)
    f = function.PrimalFunction(case.spec, np.float64, lib=lib)
    attach(f, rank, world, uid)
    out = f(*case.inputs(), replace_reusable=True)
    dtc_global = f.dtc_global()          # collective: the parallel.min(2*CFL/dtc) of the adaptive time step (adFVM/solver.py:365)
    out2 = f(*case.inputs(list(out[:3])), replace_reusable=True)
    ids = decompose.global_cell_ids(N, rank, world)
    adj = [np.ascontiguousarray(a[ids]) for a in adj_global]
    grad = f.grad()(*case.adjoint_inputs(case.state, adj))
    if tiles is not None:
        tiles[rank] = (f.tile_rounds()[2], f.tile_stats()[2])
    results[rank] = (ids, out, out2, grad, dtc_global)


def _check(world, single, results, N=N):
    g, out, out2, adj, grad = single
    for rank in range(world):
        ids, o, o2, gr = results[rank][:4]
        if len(results[rank]) > 4:
            assert relerr(results[rank][4], out[3][0, 0]) < TOL          # dtc_global: every rank holds the global maximum
        for a, b in zip(o[:3], out[:3]):
            assert relerr(a, b[ids]) < TOL
        for a, b in zip(o2[:3], out2[:3]):
            assert relerr(a, b[ids]) < TOL
        assert relerr(o[4], out[4]) < TOL                     # objective: summed over ranks
        sc = [float(np.abs(s).max()) for s in g.state]
        for grp in (slice(0, 3), slice(3, 6)):
            num = max(np.abs(a - b[ids]).max() * s for a, b, s in zip(gr[grp], grad[grp], sc))
            den = max(np.abs(b).max() * s for b, s in zip(grad[grp], sc))
            assert num / den < TOL
    # dtc: the global max is the max over ranks of the rank-local maxima (adFVM/solver.py:365 parallel.min)
    assert relerr(max(results[r][1][3][0, 0] for r in range(world)), out[3][0, 0]) < TOL


@pytest.mark.parametrize("world", [2, 4, 8])
def test_threads(world, hostsim):
    single = _single(world, hostsim)
    uid = C.create_string_buffer(128)
    hostsim.check(hostsim.dll.adfvm_comm_unique_id(uid))
    results, errors = {}, []

    def attach(f, rank, world_, uid_):
        f.c.attach_comm(uid_.raw, rank, world_)

    def run(rank):
        try:
            _rank_run(rank, world, hostsim, uid, single[3], results, attach)
        except Exception as e:      # pragma: no cover
            errors.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errors, errors
    _check(world, single, results)


@pytest.mark.parametrize("world,block", [(2, (16, 8, 8)), (4, (16, 16, 8))])
def test_threads_overlap_split(world, block, hostsim):
    """blocks large enough to have tiles that touch no processor patch: the step runs split into early tiles / cells (while
    the halo is in flight) and late ones, forward and reverse (fvm_solver.h stage / adjoint_reverse) - the only
    order-sensitive part of the multi-rank step - and must still reproduce the single-rank run"""
    single = _single(world, hostsim, block)
    uid = C.create_string_buffer(128)
    hostsim.check(hostsim.dll.adfvm_comm_unique_id(uid))
    results, tiles, errors = {}, {}, []

    def attach(f, rank, world_, uid_):
        f.c.attach_comm(uid_.raw, rank, world_)

    def run(rank):
        try:
            _rank_run(rank, world, hostsim, uid, single[3], results, attach, block, tiles)
        except Exception as e:      # pragma: no cover
            errors.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errors, errors
    for r in range(world):
        assert 0 < tiles[r][0] < tiles[r][1], tiles
    _check(world, single, results, block)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from conftest import build_hostsim
    from adfvm_b200 import _lib
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _lib.Lib(build_hostsim())
    EX = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                     C.POINTER(C.c_long), C.POINTER(C.c_long), C.c_int)
    AR = C.CFUNCTYPE(C.c_double, C.c_double, C.c_int)

    def exchange(send, recv, ncomp, npatch, peer, tag, off, cnt, sbytes):
        dt = np.float64 if sbytes == 8 else np.float32
        order = sorted(range(npatch), key=lambda i: (peer[i], tag[i]))      # both ends post shared patches in tag order
        reqs, bufs = [], []
        for i in order:
            n = cnt[i]
            s = np.ctypeslib.as_array(C.cast(send + off[i] * sbytes, C.POINTER(C.c_double if sbytes == 8 else C.c_float)), (n,))
            reqs.append(dist.isend(torch.from_numpy(s.astype(dt)), peer[i], tag=tag[i]))
            b = torch.empty(n, dtype=torch.float64 if sbytes == 8 else torch.float32)
            reqs.append(dist.irecv(b, peer[i], tag=tag[i])); bufs.append((i, b))
        for r in reqs:
            r.wait()
        for i, b in bufs:
            C.memmove(recv + off[i] * sbytes, b.numpy().ctypes.data, cnt[i] * sbytes)

    def allreduce(v, is_max):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if is_max else dist.ReduceOp.SUM)
        return float(t[0])
    ex, ar = EX(exchange), AR(allreduce)

    def attach(f, rank_, world_, uid_):
        assert lib.dll.adfvm_hostsim_comm_callback(f.c.ctx, ex, ar) == 0
    single = _single(world, lib)
    results = {}
    _rank_run(rank, world, lib, None, single[3], results, attach)
    _check_one = {rank: results[rank]}
    try:
        g, out, out2, adj, grad = single
        ids, o, o2, gr, dtc_global = results[rank]
        assert relerr(dtc_global, out[3][0, 0]) < TOL          # every rank holds the global maximum
        errs = [relerr(a, b[ids]) for a, b in zip(o2[:3], out2[:3])] + [relerr(o[4], out[4])]
        errs += [relerr(a, b[ids]) for a, b in zip(gr[:3], grad[:3])]
        q.put((rank, max(errs)))
    finally:
        dist.destroy_process_group()


def test_gloo_two_processes():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    got = [q.get(timeout=300) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert sorted(r for r, _ in got) == [0, 1]
    for _, e in got:
        assert e < 1e-9


# ---- the vane design objective (cut plane given as extraArgs) on a decomposed mesh: its mass flux is all-reduced on the
# device between the two passes, forward and in the adjoint seeds (adFVM/objectives/vane.py:97-99 mpi_allreduce)
PLANE = {"kind": "plane_ptloss", "ptin": 175158., "normal": [1., 0., 0.], "scale": 0.4, "nExtra": 5}


def _plane_extras(case, ids_global, plane_global):
    """extraArgs of a rank: the plane cells it owns (rank-local numbering) + areas; two dummy weight arrays"""
    pos = {g: i for i, g in enumerate(ids_global)}
    sel = [(pos[g], a) for g, a in plane_global if g in pos]
    cells = np.array([c for c, _ in sel], np.int32).reshape(-1, 1)
    areas = np.array([a for _, a in sel], np.float64).reshape(-1, 1)
    return [len(sel), cells, areas, np.zeros((1, 1)), np.zeros((1, 1))]


def _with_plane(case):
    spec = dict(case.spec); spec["objective"] = PLANE
    return spec


def _plane_of(g):
    """cells of the column x = 2.5 / N[0]*px of the global box with a per-cell 'area'"""
    cc = g.mesh.cellCentres[:g.mesh.nInternalCells]
    xs = np.unique(np.round(cc[:, 0], 12))
    sel = np.where(np.abs(cc[:, 0] - xs[len(xs) // 2]) < 1e-9)[0]
    return [(int(c), 1e-3 * (1 + 0.1 * np.sin(c))) for c in sel]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_threads_plane_objective(world, hostsim):
    g = decompose.global_box(N, world)
    plane = _plane_of(g)
    gids = list(range(g.mesh.nInternalCells))
    ex_g = _plane_extras(g, gids, plane)
    f = function.PrimalFunction(_with_plane(g), np.float64, lib=hostsim)
    out = f(*(g.inputs() + ex_g), replace_reusable=True)
    adj = _seed(g.state)
    a = lambda v: np.array([[v]], np.float64)
    grad = f.grad()(*(g.inputs() + ex_g + adj + [a(0.), a(1.), a(0.)]))
    uid = C.create_string_buffer(128)
    hostsim.check(hostsim.dll.adfvm_comm_unique_id(uid))
    results, errors = {}, []

    def run(rank):
        try:
            case = decompose.periodic_box_rank(N, rank, world)
            ids = decompose.global_cell_ids(N, rank, world)
            ex_r = _plane_extras(case, list(ids), plane)
            fr = function.PrimalFunction(_with_plane(case), np.float64, lib=hostsim)
            fr.c.attach_comm(uid.raw, rank, world)
            o = fr(*(case.inputs() + ex_r), replace_reusable=True)
            adj_r = [np.ascontiguousarray(x[ids]) for x in adj]
            gr = fr.grad()(*(case.inputs() + ex_r + adj_r + [a(0.), a(1.), a(0.)]))
            results[rank] = (ids, o, gr)
        except Exception as e:      # pragma: no cover
            errors.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errors, errors
    assert abs(out[4][0, 0]) > 1e-6                      # a real objective value
    sc = [float(np.abs(s).max()) for s in g.state]
    for rank in range(world):
        ids, o, gr = results[rank]
        assert relerr(o[4], out[4]) < TOL
        for grp in (slice(0, 3), slice(3, 6)):
            num = max(np.abs(x - y[ids]).max() * s for x, y, s in zip(gr[grp], grad[grp], sc))
            den = max(np.abs(y).max() * s for y, s in zip(grad[grp], sc))
            assert num / den < TOL


# ---- a general (non-periodic, walled) mesh decomposed like OpenFOAM's decomposePar: physical patches split between the
# ranks (some left without faces), processor patches from the cut faces, flipped faces on the neighbour side
@pytest.mark.parametrize("world", [2, 4, 8])
def test_threads_decomposed_walled_mesh(world, hostsim):
    from adfvm_b200 import cases
    g = cases.walled_box((8, 6, 4), warp=0.0)            # planar faces: decomposition invariance is exact (see decompose.py)
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim)
    out = f(*g.inputs(), replace_reusable=True)
    out2 = f(*g.inputs(list(out[:3])), replace_reusable=True)
    adj = _seed(g.state)
    grad = f.grad()(*g.adjoint_inputs(g.state, adj))
    parts = decompose.rank_cases(g, world)
    assert any(p[0].mesh.boundary["inlet"]["nFaces"] == 0 for p in parts)          # a rank without inlet faces
    uid = C.create_string_buffer(128)
    hostsim.check(hostsim.dll.adfvm_comm_unique_id(uid))
    results, errors = {}, []

    def run(rank):
        try:
            case, ids = parts[rank]
            fr = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
            fr.c.attach_comm(uid.raw, rank, world)
            o = fr(*case.inputs(), replace_reusable=True)
            o2 = fr(*case.inputs(list(o[:3])), replace_reusable=True)
            gr = fr.grad()(*case.adjoint_inputs(case.state, [np.ascontiguousarray(x[ids]) for x in adj]))
            results[rank] = (ids, o, o2, gr)
        except Exception as e:      # pragma: no cover
            errors.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errors, errors
    _check(world, (g, out, out2, adj, grad), results)


# ---- loading a decomposed case from disk the way every MPI rank of the reference does (processor<r>/constant/polyMesh in
# OpenFOAM binary + exchange of the cell centres across processor patches), two real processes over gloo
def _load_worker(rank, world, port, case_dir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from adfvm_b200 import cases
    from adfvm_b200.metrics import GRAD_FIELDS, INT_FIELDS
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = decompose.load_decomposed_mesh(case_dir, rank)
        ref = decompose.rank_cases(cases.walled_box((8, 6, 4), warp=0.0), world)[rank][0].mesh
        err = 0.
        for a in GRAD_FIELDS:
            err = max(err, float(np.abs(getattr(m, a) - getattr(ref, a)).max() / np.abs(getattr(ref, a)).max()))
        ok = all(np.array_equal(getattr(m, a), getattr(ref, a)) for a in INT_FIELDS) and m.getScalar() == ref.getScalar()
        q.put((rank, err, ok))
    finally:
        dist.destroy_process_group()


def test_load_decomposed_case_two_processes(tmp_path):
    import torch.multiprocessing as mp
    from adfvm_b200 import cases, hexmesh
    g = cases.walled_box((8, 6, 4), warp=0.0).mesh
    poly = hexmesh.PolyMesh(g.points, g.faces, g.owner, g.neighbour[:g.nInternalFaces],
                            {k: {kk: vv for kk, vv in v.items() if kk != "cellStartFace"} for k, v in g.boundary.items()})
    parts = decompose.decompose_polymesh(poly, decompose.slab_partition(g.cellCentres[:g.nInternalCells], 2))
    decompose.write_decomposed_case(str(tmp_path), parts)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_load_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    [p.start() for p in ps]
    got = [q.get(timeout=300) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert sorted(r for r, _, _ in got) == [0, 1]
    for _, err, ok in got:
        assert ok and err < 1e-12


# ---- a fully periodic mesh through the general decomposer: the partition separates the two cells of every x-cyclic pair,
# which turns those faces into processorCyclic patches next to the plain processor patch of the cut
@pytest.mark.parametrize("world", [2, 3])
def test_threads_decomposed_periodic_mesh(world, hostsim):
    from adfvm_b200 import cases, hexmesh
    from adfvm_b200.metrics import build_mesh
    g = cases.periodic_box((6, 5, 4), mesh=build_mesh(hexmesh.box_mesh((6, 5, 4))))
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim)
    out = f(*g.inputs(), replace_reusable=True)
    out2 = f(*g.inputs(list(out[:3])), replace_reusable=True)
    adj = _seed(g.state)
    grad = f.grad()(*g.adjoint_inputs(g.state, adj))
    parts = decompose.rank_cases(g, world)
    types = {p["type"] for c, _ in parts for p in c.mesh.boundary.values()}
    assert "processorCyclic" in types and "processor" in types
    uid = C.create_string_buffer(128)
    hostsim.check(hostsim.dll.adfvm_comm_unique_id(uid))
    results, errors = {}, []

    def run(rank):
        try:
            case, ids = parts[rank]
            fr = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
            fr.c.attach_comm(uid.raw, rank, world)
            o = fr(*case.inputs(), replace_reusable=True)
            o2 = fr(*case.inputs(list(o[:3])), replace_reusable=True)
            gr = fr.grad()(*case.adjoint_inputs(case.state, [np.ascontiguousarray(x[ids]) for x in adj]))
            results[rank] = (ids, o, o2, gr)
        except Exception as e:      # pragma: no cover
            errors.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errors, errors
    _check(world, (g, out, out2, adj, grad), results)


# ---- parameters = 'mesh' on a decomposed mesh: every rank returns the gradients with respect to ITS OWN metric arrays (the
# reference's ranks hold their own mesh files); the sensitivity to a perturbation of the points, summed over the ranks
# (cmesh.computeSensitivity + parallel.sum, apps/adjoint.py:337-341), must equal the single-rank one
@pytest.mark.parametrize("world", [2, 4])
def test_threads_mesh_sensitivity(world, hostsim):
    from adfvm_b200 import cases, hexmesh
    from adfvm_b200.metrics import build_mesh, GRAD_FIELDS
    g = cases.walled_box((8, 6, 4), warp=0.0)
    g.spec = dict(g.spec, parameters="mesh")
    adj = _seed(g.state)
    grads = function.PrimalFunction(g.spec, np.float64, lib=hostsim).grad()(*g.adjoint_inputs(g.state, adj))[3:]
    parts = decompose.rank_cases(g, world)
    gm = g.mesh

    def poly_of(mesh, pts):
        return hexmesh.PolyMesh(pts, mesh.faces, mesh.owner, mesh.neighbour[:mesh.nInternalFaces],
                                {k: {kk: vv for kk, vv in v.items() if kk != "cellStartFace"} for k, v in mesh.boundary.items()})
    # directional derivative of every metric array along a smooth displacement of the points (central difference)
    eps = 1e-6
    disp = np.stack([0.3 * np.sin(2 * gm.points[:, 1]), 0.2 * gm.points[:, 0] * gm.points[:, 2], 0.1 * np.cos(gm.points[:, 0])], axis=1)
    mp, mm = build_mesh(poly_of(gm, gm.points + eps * disp)), build_mesh(poly_of(gm, gm.points - eps * disp))
    S_single = sum(float((np.asarray(gr).reshape(getattr(mp, a).shape) * (getattr(mp, a) - getattr(mm, a)) / (2 * eps)).sum())
                   for a, gr in zip(GRAD_FIELDS, grads))
    uid = C.create_string_buffer(128)
    hostsim.check(hostsim.dll.adfvm_comm_unique_id(uid))
    results, errors = {}, []

    def run(rank):
        try:
            case, ids = parts[rank]
            case.spec = dict(case.spec, parameters="mesh")
            fr = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
            fr.c.attach_comm(uid.raw, rank, world)
            results[rank] = fr.grad()(*case.adjoint_inputs(case.state, [np.ascontiguousarray(x[ids]) for x in adj]))[3:]
        except Exception as e:      # pragma: no cover
            errors.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    assert not errors, errors
    # the ranks' meshes rebuilt from the displaced points (same partition, ghost centres from the displaced global mesh)
    S_ranks = 0.
    for sign, gmesh in ((+1, mp), (-1, mm)):
        gmesh.boundary = {k: dict(v) for k, v in gmesh.boundary.items()}
        pr = decompose.decompose_polymesh(poly_of(gmesh, gmesh.points), decompose.slab_partition(gm.cellCentres[:gm.nInternalCells], world))
        for r in range(world):
            mr = build_mesh(pr[r]["poly"], decompose.remote_centres(pr, r, gmesh))
            S_ranks += sign * sum(float((np.asarray(gr).reshape(getattr(mr, a).shape) * getattr(mr, a)).sum())
                                  for a, gr in zip(GRAD_FIELDS, results[r])) / (2 * eps)
    assert abs(S_single) > 0
    assert abs(S_ranks - S_single) <= 1e-6 * abs(S_single), (S_ranks, S_single)


# ---- BASELINE.json config 4: the vane cascade (blockMeshDict of cases/vane_optim, spline edges, 3-D: several spanwise layers) with
# the reference's design objective (cut plane as extraArgs, its mass flux all-reduced) decomposed like decomposePar
@pytest.mark.parametrize("world", [2, 4])
def test_threads_vane_cascade(world, hostsim):
    from adfvm_b200 import cases
    g = cases.vane_cascade(nz=2)
    f = function.PrimalFunction(g.spec, np.float64, lib=hostsim)
    out = f(*g.inputs(), replace_reusable=True)
    adj = _seed(g.state)
    grad = f.grad()(*g.adjoint_inputs(g.state, adj))
    assert abs(out[4][0, 0]) > 1e-3
    parts = decompose.rank_cases(g, world)
    assert sum(p[0].extra[0] for p in parts) == g.extra[0] and sum(1 for p in parts if p[0].extra[0] > 0) >= 1
    uid = C.create_string_buffer(128)
    hostsim.check(hostsim.dll.adfvm_comm_unique_id(uid))
    results, errors = {}, []

    def run(rank):
        try:
            case, ids = parts[rank]
            fr = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
            fr.c.attach_comm(uid.raw, rank, world)
            o = fr(*case.inputs(), replace_reusable=True)
            gr = fr.grad()(*case.adjoint_inputs(case.state, [np.ascontiguousarray(x[ids]) for x in adj]))
            results[rank] = (ids, o, gr)
        except Exception as e:      # pragma: no cover
            errors.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=600) for t in ts]
    assert not errors, errors
    sc = [float(np.abs(s).max()) for s in g.state]
    for rank in range(world):
        ids, o, gr = results[rank]
        for a, b in zip(o[:3], out[:3]):
            assert relerr(a, b[ids]) < TOL
        assert relerr(o[4], out[4]) < TOL
        for grp in (slice(0, 3), slice(3, 6)):
            num = max(np.abs(a - b[ids]).max() * s for a, b, s in zip(gr[grp], grad[grp], sc))
            den = max(np.abs(b).max() * s for b, s in zip(grad[grp], sc))
            assert num / den < TOL
