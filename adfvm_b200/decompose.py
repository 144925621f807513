"""Rank-local pieces of a block-decomposed periodic hex box, in the reference's decomposed-case conventions
(`processorN/constant/polyMesh/boundary`, reference adFVM/mesh.py:238-249, 746-756, 784-805):

* every rank owns an (nx,ny,nz) block of a (nx*px, ny*py, nz*pz) periodic box (weak scaling: fixed work per rank);
* a block side whose direction is not split (p == 1) stays a local `cyclic` pair, exactly as on one rank;
* a side in a split direction becomes a `processor` patch (interior cut) or a `processorCyclic` patch (the cut
  coincides with the periodic wrap), named `procBoundary<i>to<j>[through<cyc>]` like OpenFOAM's decomposePar,
  listed after all physical patches so that remote ghost rows are the tail of the cell arrays, with
  `myProcNo/neighbProcNo` and a `tag` shared by both sides; face i of a patch matches face i of the peer's patch.

The halo itself (pack -> NCCL send/recv -> unpack, reverse + add for the adjoint) lives in the native library
(adfvm_comm_init, replacing adFVM/cpp/parallel.cpp:31-209); `attach_comm` distributes the communicator id over
torch.distributed, which is plumbing only.
"""
from __future__ import annotations

import numpy as np

from . import cases
from .metrics import uniform_box


def factor3(world):
    """(px, py, pz) with px >= py >= pz, as cubic as possible: 1->(1,1,1) 2->(2,1,1) 4->(2,2,1) 8->(2,2,2)"""
    best = None
    for px in range(1, world + 1):
        if world % px:
            continue
        for py in range(1, world // px + 1):
            if (world // px) % py:
                continue
            pz = world // px // py
            if px >= py >= pz:
                key = (px - pz, px)
                if best is None or key < best[0]:
                    best = (key, (px, py, pz))
    return best[1]


def rank_coords(rank, p):
    return rank % p[0], (rank // p[0]) % p[1], rank // (p[0] * p[1])


def coords_rank(c, p):
    return (c[0] % p[0]) + p[0] * ((c[1] % p[1]) + p[1] * (c[2] % p[2]))


def _side_tag(dim, plus, me, peer):
    # identical on both ends of a connection: my '+' side meets the peer's '-' side
    return 2 * dim + (1 if plus == (me < peer) else 0)


def periodic_box_rank(n, rank, world, dtype=np.float64, p=None, dt=None):
    """Case of `rank` out of `world` (see module docstring). n: cells per side per rank (int or 3-tuple)."""
    if isinstance(n, int):
        n = (n, n, n)
    p = p or factor3(world)
    assert p[0] * p[1] * p[2] == world
    me = rank_coords(rank, p)
    L = tuple(1.0 for _ in range(3))                      # block edge length; global box = p[d] * L[d]
    lo = tuple(me[d] * L[d] for d in range(3))
    hi = tuple(lo[d] + L[d] for d in range(3))
    phys, proc = [], []
    names = "xyz"
    for d in range(3):
        c1, c2 = names[d] + "1", names[d] + "2"
        if p[d] == 1:
            phys.append((c1, "cyclic", [names[d] + "-"], {"neighbourPatch": c2}))
            phys.append((c2, "cyclic", [names[d] + "+"], {"neighbourPatch": c1}))
            continue
        for plus in (False, True):
            nb = list(me); nb[d] += 1 if plus else -1
            wrap = nb[d] < 0 or nb[d] >= p[d]
            peer = coords_rank(nb, p)
            extra = {"myProcNo": rank, "neighbProcNo": peer, "tag": _side_tag(d, plus, rank, peer)}
            name = "procBoundary%dto%d" % (rank, peer)
            ptype = "processor"
            if wrap:
                ptype = "processorCyclic"
                extra["referPatch"] = c2 if plus else c1
                name += "through" + extra["referPatch"]
            proc.append((name, ptype, [names[d] + ("+" if plus else "-")], extra))
    mesh = uniform_box(n, lo, hi, phys + proc)            # ghost centres of processor faces = the peer block's cells
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    U, T, pr = cases.smooth_state(cc)                     # period 1 in every direction: periodic on the global box too
    bcs = {f: {pid: {"type": "cyclic", "keys": []} for pid in mesh.sortedPatches} for f in ("U", "T", "p")}
    spec = cases._spec(mesh, bcs, {"kind": "cell_TV"})
    if dt is None:
        dt = 1e-6 * 48. / max(n)
    mid = tuple(0.5 * p[d] * L[d] for d in range(3))
    return cases.Case(mesh, spec, cases.conservative(U, T, pr), cases.gaussian_source(cc, mid), {}, dt, dtype)


def global_box(n, world, dtype=np.float64, p=None, dt=None):
    """The undecomposed mesh of the same problem (single rank), for decomposition-invariance checks
    (the reference's own criterion, tests/test_parallel.py:63-81)."""
    if isinstance(n, int):
        n = (n, n, n)
    p = p or factor3(world)
    N = tuple(n[d] * p[d] for d in range(3))
    hi = tuple(float(p[d]) for d in range(3))
    mesh = uniform_box(N, (0., 0., 0.), hi)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    U, T, pr = cases.smooth_state(cc)
    bcs = {f: {pid: {"type": "cyclic", "keys": []} for pid in mesh.sortedPatches} for f in ("U", "T", "p")}
    spec = cases._spec(mesh, bcs, {"kind": "cell_TV"})
    if dt is None:
        dt = 1e-6 * 48. / max(n)
    mid = tuple(0.5 * hi[d] for d in range(3))
    return cases.Case(mesh, spec, cases.conservative(U, T, pr), cases.gaussian_source(cc, mid), {}, dt, dtype)


def global_cell_ids(n, rank, world, p=None):
    """global cell index (in `global_box` numbering) of every cell of `rank`'s block"""
    if isinstance(n, int):
        n = (n, n, n)
    p = p or factor3(world)
    me = rank_coords(rank, p)
    K, J, I = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
    gi, gj, gk = I.ravel() + me[0] * n[0], J.ravel() + me[1] * n[1], K.ravel() + me[2] * n[2]
    return gi + n[0] * p[0] * (gj + n[1] * p[1] * gk)


def attach_comm(f, rank, world, group=None):
    """Create the native halo communicator of PrimalFunction `f`: rank 0 makes the id, torch.distributed
    (already initialised by the caller: nccl on GPUs) broadcasts it."""
    import ctypes
    import torch
    import torch.distributed as dist
    lib = f.c.lib
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        lib.check(lib.dll.adfvm_comm_unique_id(buf))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0, group=group)
    f.c.attach_comm(bytes(t.cpu().tolist()), rank, world)


# ------------------------------------------------------------------------------------------ general meshes
def slab_partition(cell_centres, nranks, axis=0):
    """cell -> rank by equal-count slabs along `axis` (OpenFOAM decomposePar `simple` with n = (nranks 1 1))"""
    order = np.lexsort((np.arange(len(cell_centres)), cell_centres[:, axis]))
    rank = np.empty(len(cell_centres), np.int64)
    rank[order] = (np.arange(len(order)) * nranks) // len(order)
    return rank


def decompose_polymesh(poly, cell_rank):
    """Split a `hexmesh.PolyMesh` into rank-local meshes in the layout of OpenFOAM's decomposePar that the reference reads
    (`processor<r>/constant/polyMesh`, adFVM/mesh.py:177-204, 238-249): local cells in ascending global order, the
    rank's inner faces first, then EVERY physical patch (possibly with 0 faces, same order and attributes), then the
    processor patches: `procBoundary<r>to<q>` with the cut faces in ascending global face order on both sides (face i of
    r's patch is face i of q's), flipped on the side that held the neighbour cell so that normals leave the local owner;
    and, where the partition separates the two cells of a cyclic pair, `procBoundary<r>to<q>through<cyc>` (type
    processorCyclic, referPatch = the cyclic patch the local faces came from) in ascending pair order. Tags are equal on
    the two ends of a connection and distinct between the connections of one pair of ranks (adFVM/mesh.py:746-756).
    Returns per rank a dict(poly, cellProcAddressing, faceProcAddressing [global face of every local face], patchFaces
    {patch: positions of the local faces inside the global patch})."""
    from collections import OrderedDict
    from .hexmesh import PolyMesh
    cell_rank = np.asarray(cell_rank, np.int64)
    nranks = int(cell_rank.max()) + 1
    owner = np.asarray(poly.owner, np.int64); neigh = np.asarray(poly.neighbour, np.int64)
    nIF = len(neigh)
    names = list(poly.boundary)
    # rank of the cell behind every boundary face of a cyclic patch (its partner face's owner), else -1
    partner_rank = {}
    for name, p in poly.boundary.items():
        if p["type"] == "cyclic":
            q = poly.boundary[p["neighbourPatch"]]
            partner_rank[name] = cell_rank[owner[q["startFace"]:q["startFace"] + q["nFaces"]]]
    ro, rn = cell_rank[owner[:nIF]], cell_rank[neigh]
    out = []
    for r in range(nranks):
        cells = np.where(cell_rank == r)[0]
        g2l = np.full(len(cell_rank), -1, np.int64); g2l[cells] = np.arange(len(cells))
        inner = np.where((ro == r) & (rn == r))[0]
        f_list, own_l, faces_l = [inner], [g2l[owner[inner]]], [poly.faces[inner]]
        nb_l = g2l[neigh[inner]]
        boundary = OrderedDict()
        patch_faces = {}
        proc = []                                  # (peer, tag, name, dict, global faces, local owners, face vertices)
        start = len(inner)
        for name, p in poly.boundary.items():
            gf = np.arange(p["startFace"], p["startFace"] + p["nFaces"])
            mine = cell_rank[owner[gf]] == r
            if name in partner_rank:               # faces whose periodic partner lives on another rank leave the cyclic patch
                pr = partner_rank[name]
                away = mine & (pr != r)
                mine = mine & (pr == r)
                ia, ib = names.index(name), names.index(p["neighbourPatch"])
                for q in sorted(set(pr[away].tolist())):
                    sel = gf[away & (pr == q)]
                    tag = 1 + 2 * min(ia, ib) + (1 if ((r < q) == (ia < ib)) else 0)
                    proc.append((q, tag, "procBoundary%dto%dthrough%s" % (r, q, name),
                                 OrderedDict(type="processorCyclic", myProcNo=int(r), neighbProcNo=int(q), referPatch=name, tag=int(tag)),
                                 sel, g2l[owner[sel]], poly.faces[sel]))
            sel = gf[mine]
            d = OrderedDict((k, v) for k, v in p.items() if k not in ("nFaces", "startFace", "cellStartFace"))
            d["nFaces"], d["startFace"] = int(len(sel)), int(start)
            boundary[name] = d
            patch_faces[name] = sel - p["startFace"]
            f_list.append(sel); own_l.append(g2l[owner[sel]]); faces_l.append(poly.faces[sel])
            start += len(sel)
        cut = np.where((ro == r) != (rn == r))[0]                 # global inner faces with exactly one cell here
        peers = np.where(ro[cut] == r, rn[cut], ro[cut])
        for q in sorted(set(peers.tolist())):
            sel = cut[peers == q]                                  # ascending global face order on both ranks
            mine_is_owner = ro[sel] == r
            fl = poly.faces[sel].copy()
            fl[~mine_is_owner] = fl[~mine_is_owner][:, [0, 3, 2, 1]]   # normal must leave the local cell (OpenFOAM reverseFace:
            # first vertex kept). NB the reference takes the normal from a face's first three points (cmesh.cpp:72-87): on
            # non-planar faces a reversed face has a slightly different normal, so decomposition invariance is exact only
            # for meshes with planar faces - with decomposePar and the reference as well
            lo = np.where(mine_is_owner, g2l[owner[sel]], g2l[neigh[sel]])
            proc.append((q, 0, "procBoundary%dto%d" % (r, q),
                         OrderedDict(type="processor", myProcNo=int(r), neighbProcNo=int(q), tag=0), sel, lo, fl))
        for q, tag, name, d, sel, lo, fl in sorted(proc, key=lambda x: (x[0], x[1])):
            d["nFaces"], d["startFace"] = int(len(sel)), int(start)
            boundary[name] = d
            f_list.append(sel); own_l.append(lo); faces_l.append(fl)
            start += len(sel)
        out.append(dict(poly=PolyMesh(poly.points, np.concatenate(faces_l), np.concatenate(own_l), nb_l, boundary),
                        cellProcAddressing=cells, faceProcAddressing=np.concatenate(f_list), patchFaces=patch_faces))
    return out


def remote_centres(parts, rank, global_mesh):
    """ghost-cell centres of `rank`'s processor patches = centres of the peer's cells across the cut faces, shifted by the
    periodic offset for processorCyclic patches (what Mesh.createGhostCells exchanges over MPI, adFVM/mesh.py:784-805);
    here taken from the undecomposed mesh `global_mesh` (metrics.MeshData)"""
    part = parts[rank]
    gm = global_mesh
    o_all, n_all = np.asarray(gm.owner, np.int64), np.asarray(gm.neighbour, np.int64)
    mine = np.zeros(gm.nInternalCells, bool); mine[part["cellProcAddressing"]] = True
    res = {}
    for name, p in part["poly"].boundary.items():
        gf = part["faceProcAddressing"][p["startFace"]:p["startFace"] + p["nFaces"]]
        if p["type"] == "processor":
            o, n = o_all[gf], n_all[gf]
            res[name] = gm.cellCentres[np.where(mine[o], n, o)]
        elif p["type"] == "processorCyclic":
            ref = gm.boundary[p["referPatch"]]
            nbp = gm.boundary[ref["neighbourPatch"]]
            pf = nbp["startFace"] + (gf - ref["startFace"])                        # partner faces
            res[name] = gm.cellCentres[o_all[pf]] + (gm.faceCentres[gf] - gm.faceCentres[pf])
    return res


def rank_cases(case, nranks, axis=0):
    """Decompose a single-rank `cases.Case` (any PolyMesh-backed mesh with planar faces) into `nranks` rank-local cases:
    slab partition -> decompose_polymesh -> metrics with the peers' cell centres -> state / source / BC arrays sliced by
    cellProcAddressing and by the patch face maps. Returns [(Case, cellProcAddressing)]."""
    from .hexmesh import PolyMesh
    from .metrics import build_mesh
    gm = case.mesh
    C = gm.nInternalCells
    poly = PolyMesh(gm.points, gm.faces, gm.owner, gm.neighbour[:gm.nInternalFaces],
                    {k: {kk: vv for kk, vv in v.items() if kk != "cellStartFace"} for k, v in gm.boundary.items()})
    parts = decompose_polymesh(poly, slab_partition(gm.cellCentres[:C], nranks, axis))
    out = []
    for r, part in enumerate(parts):
        rc = remote_centres(parts, r, gm)
        m = build_mesh(part["poly"], rc)
        ids = part["cellProcAddressing"]
        spec = dict(case.spec)
        spec["patches"] = cases._spec(m, spec["BCs"], spec.get("objective"))["patches"]
        spec["sortedPatches"] = list(m.sortedPatches)
        bcvals = {(f, pid, key): v[part["patchFaces"][pid]] for (f, pid, key), v in case.bcvals.items()}
        rcase = cases.Case(m, spec, [s[ids] for s in case.state], [s[ids] for s in case.source], bcvals, case.dt, case.dtype)
        extras = getattr(case, "extra", None)
        if extras is not None:
            # extraArgs of the cut-plane objective (adFVM/objectives/vane.py): the plane cells owned by this rank with their areas,
            # then one weight array per listed physical patch, restricted to the rank's faces of it
            pos = -np.ones(C, np.int64); pos[ids] = np.arange(len(ids))
            gcells = np.asarray(extras[1]).ravel()
            mine = pos[gcells] >= 0
            rcase.extra = [int(mine.sum()), np.ascontiguousarray(pos[gcells][mine].astype(np.int32).reshape(-1, 1)),
                           np.ascontiguousarray(np.asarray(extras[2])[mine])]
            for w, pid in zip(extras[3:], getattr(case, "extra_patches", [])):
                rcase.extra.append(np.ascontiguousarray(np.asarray(w)[part["patchFaces"][pid]]))
        out.append((rcase, ids))
    return out


def write_decomposed_case(case_dir, parts):
    """processor<r>/constant/polyMesh/* of every part (OpenFOAM binary, adfvm_b200.foam_io) + the addressing arrays"""
    import os
    from . import foam_io
    for r, part in enumerate(parts):
        d = os.path.join(case_dir, "processor%d" % r)
        foam_io.write_polymesh(d, part["poly"])
        np.save(os.path.join(d, "constant", "polyMesh", "cellProcAddressing.npy"), part["cellProcAddressing"])


def load_decomposed_mesh(case_dir, rank, group=None):
    """Rank-local mesh of an OpenFOAM-decomposed case as the reference builds it in every MPI rank
    (Mesh.readFoam + createGhostCells, adFVM/mesh.py:177-204, 758-819): read `processor<rank>/constant/polyMesh`, compute
    the cell centres of the own cells, exchange those next to processor patches with the peer ranks (torch.distributed
    point-to-point, any backend) and build the metrics with them."""
    import os
    import torch
    import torch.distributed as dist
    from . import foam_io
    from .metrics import build_mesh
    poly = foam_io.read_polymesh(os.path.join(case_dir, "processor%d" % rank))
    proc = [(k, v) for k, v in poly.boundary.items() if v["type"] in ("processor", "processorCyclic")]
    for k, v in proc:
        v.setdefault("tag", 0)
    own = build_mesh(poly, {k: np.zeros((v["nFaces"], 3)) for k, v in proc})    # own-cell centres do not depend on ghosts
    cc = own.cellCentres[:own.nInternalCells]
    reqs, recv = [], {}
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    for k, v in sorted(proc, key=lambda kv: (kv[1]["neighbProcNo"], int(kv[1].get("tag", 0)))):
        s, n = v["startFace"], v["nFaces"]
        send = torch.from_numpy(np.ascontiguousarray(cc[np.asarray(poly.owner[s:s + n], np.int64)])).to(dev)
        buf = torch.empty((n, 3), dtype=torch.float64, device=dev)
        recv[k] = buf
        reqs.append(dist.isend(send, int(v["neighbProcNo"]), group=group, tag=int(v.get("tag", 0))))
        reqs.append(dist.irecv(buf, int(v["neighbProcNo"]), group=group, tag=int(v.get("tag", 0))))
    for q in reqs:
        q.wait()
    return build_mesh(poly, {k: b.cpu().numpy() for k, b in recv.items()})
