"""Mesh connectivity + metric arrays consumed by the residual hot path.

Restates, in vectorised numpy, what the reference computes once per mesh in
`adFVM/cpp/cmesh.cpp:3-270` (+ `adFVM/mesh.py:84-127, 758-819`): the 10 `Mesh.gradFields`
(areas, volumesL, volumesR, weights, deltas, normals, deltasUnit, linearWeights,
quadraticWeights, volumes) and the 5 `Mesh.intFields` (owner, neighbour, cellFaces,
cellNeighbours, cellOwner), plus the 8 size constants and the per-patch
(startFace, nFaces, cellStartFace) triples (reference `adFVM/mesh.py:27-37, 870-881`).

This is host-side setup (stays on the CPU); it runs once per mesh.
The arithmetic follows the reference's order of operations so that small meshes agree
with `cmesh.build` to round-off (checked in tests/test_metrics.py against golden fixtures).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

PROCESSOR_PATCHES = ("processor", "processorCyclic")
CYCLIC_PATCHES = ("cyclic", "slidingPeriodic1D")
COUPLED_PATCHES = CYCLIC_PATCHES + PROCESSOR_PATCHES

GRAD_FIELDS = ["areas", "volumesL", "volumesR", "weights", "deltas", "normals", "deltasUnit",
               "linearWeights", "quadraticWeights", "volumes"]
INT_FIELDS = ["owner", "neighbour", "cellFaces", "cellNeighbours", "cellOwner"]
CONSTANTS = ["nCells", "nFaces", "nInternalCells", "nInternalFaces",
             "nLocalCells", "nRemoteCells", "nLocalFaces", "nGhostCells"]
BC_FIELDS = ["startFace", "nFaces", "cellStartFace"]


def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                     a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


class MeshData:
    """All arrays of one (rank-local) mesh, named as in the reference `Mesh` class."""

    def getTensor(self):
        """reference adFVM/mesh.py:870-872"""
        return [getattr(self, a) for a in GRAD_FIELDS] + [getattr(self, a) for a in INT_FIELDS]

    def getScalar(self):
        """reference adFVM/mesh.py:874-881"""
        out = [int(getattr(self, a)) for a in CONSTANTS]
        for patchID in self.sortedPatches:
            out.extend(int(self.boundary[patchID][a]) for a in BC_FIELDS)
        return out

    def astype(self, dtype):
        """Copy with scalar arrays cast (fp32 runs: reference casts points first, mesh.py:191;
        we build in fp64 and round once, which is at least as accurate)."""
        import copy
        m = copy.copy(self)
        for a in GRAD_FIELDS + ["cellCentres", "faceCentres"]:
            setattr(m, a, np.ascontiguousarray(getattr(self, a), dtype))
        return m


def build_mesh(poly, remote_centres=None, device=None):
    """poly: hexmesh.PolyMesh (rank-local). remote_centres: for processor patches, dict
    patchID -> [nFaces,3] ghost cell centres (what the reference exchanges over MPI in
    `createGhostCells`, mesh.py:784-805).
    device: None -> the numpy restatement below (test infrastructure for synthetic inputs); a dict
    {"lib": _lib.Lib, "device": 0, "precision": np.float64, "stream": None} -> the geometry (everything of
    adFVM/cpp/cmesh.cpp:65-233 and the ghost-cell centres) is computed by the library on the device
    (adfvm_mesh_metrics, csrc/fvm_metrics.h); the integer connectivity stays here."""
    if device is not None:
        return _build_mesh_device(poly, remote_centres, device)
    m = MeshData()
    points, faces = poly.points, poly.faces.astype(np.int64)
    owner = poly.owner.astype(np.int64)
    nb_int = poly.neighbour.astype(np.int64)
    nF, nIF = len(owner), len(nb_int)
    nIC = int(owner.max()) + 1
    nG = nF - nIF
    m.boundary = OrderedDict((k, dict(v)) for k, v in poly.boundary.items())
    m.nFaces, m.nInternalFaces, m.nInternalCells = nF, nIF, nIC
    m.nGhostCells, m.nCells = nG, nIC + nG
    m.nBoundaryFaces = nG

    # --- cellFaces: owned faces ascending, then faces where the cell is neighbour (cmesh.cpp:9-19)
    cells_of_face = np.concatenate([owner, nb_int])
    face_ids = np.concatenate([np.arange(nF), np.arange(nIF)])
    pri = np.concatenate([np.zeros(nF, np.int64), np.ones(nIF, np.int64)])
    order = np.lexsort((face_ids, pri, cells_of_face))
    assert len(order) == 6 * nIC, "hexahedral cells only (6 faces per cell)"
    cellFaces = face_ids[order].reshape(nIC, 6)
    assert np.all(cells_of_face[order].reshape(nIC, 6) == np.arange(nIC)[:, None])

    # --- face normals, centres, areas (cmesh.cpp:69-121)
    P = points[faces]                      # [nF,4,3]
    a, b, c = P[:, 0], P[:, 1], P[:, 2]
    nrm = _cross(a - b, b - c)
    nrm = nrm / np.sqrt(nrm[:, 0] * nrm[:, 0] + nrm[:, 1] * nrm[:, 1] + nrm[:, 2] * nrm[:, 2])[:, None]
    fc0 = (((P[:, 0] + P[:, 1]) + P[:, 2]) + P[:, 3]) / 4
    area = np.zeros(nF)
    sumC = np.zeros((nF, 3))
    for j in range(4):
        p0, p1 = P[:, j], P[:, (j + 1) % 4]
        avg = (p0 + p1 + fc0) / 3
        N = _cross(p1 - p0, fc0 - p0)
        Ns = np.sqrt(N[:, 0] * N[:, 0] + N[:, 1] * N[:, 1] + N[:, 2] * N[:, 2])
        area = area + Ns / 2
        sumC = sumC + (Ns[:, None] * avg) / 2
    faceCentres = sumC / area[:, None]

    # --- cell centres, volumes (cmesh.cpp:126-162)
    cc0 = np.zeros((nIC, 3))
    for j in range(6):
        cc0 = cc0 + faceCentres[cellFaces[:, j]]
    cc0 = cc0 / 6
    vol = np.zeros(nIC)
    sumCC = np.zeros((nIC, 3))
    for j in range(6):
        f = cellFaces[:, j]
        height = cc0 - faceCentres[f]
        areaN = area[f][:, None] * nrm[f]
        v = (areaN[:, 0] * height[:, 0] + areaN[:, 1] * height[:, 1]) + areaN[:, 2] * height[:, 2]
        v = np.abs(v / 3)
        avgC = 3. / 4 * faceCentres[f] + 1. / 4 * cc0
        vol = vol + v
        sumCC = sumCC + v[:, None] * avgC
    cellCentres_int = sumCC / vol[:, None]

    # --- ghost cells (mesh.py:758-819)
    neighbour = np.concatenate([nb_int, np.zeros(nG, np.int64)])
    cellCentres = np.concatenate([cellCentres_int, np.zeros((nG, 3))])
    delta_cf = nIC - nIF
    local, remote = [], []
    pids = sorted(m.boundary.keys(), key=lambda x: (m.boundary[x]["startFace"], m.boundary[x]["nFaces"]))
    nLocalCells = nIC
    for pid in pids:
        patch = m.boundary[pid]
        patch["nFaces"] = int(patch["nFaces"]); patch["startFace"] = int(patch["startFace"])
        patch["cellStartFace"] = patch["startFace"] + delta_cf
        (remote if patch["type"] in PROCESSOR_PATCHES else local).append(pid)
    for pid in pids:
        patch = m.boundary[pid]
        s, n = patch["startFace"], patch["nFaces"]
        if n == 0:
            continue
        cs = patch["cellStartFace"]
        if patch["type"] not in PROCESSOR_PATCHES:
            nLocalCells += n
        neighbour[s:s + n] = np.arange(cs, cs + n)
        if patch["type"] in CYCLIC_PATCHES:
            nbp = m.boundary[patch["neighbourPatch"]]
            ns = nbp["startFace"]
            idx = owner[ns:ns + n]
            transform = faceCentres[s] - faceCentres[ns]
            cellCentres[cs:cs + n] = transform + cellCentres[idx]
        elif patch["type"] in PROCESSOR_PATCHES:
            assert remote_centres is not None and pid in remote_centres, \
                "processor patch %s needs the neighbour rank's cell centres" % pid
            cellCentres[cs:cs + n] = remote_centres[pid]
        else:
            cellCentres[cs:cs + n] = faceCentres[s:s + n]
    # processor patches must be the tail so remote ghost rows are [nLocalCells, nCells) (mesh.py:110-112)
    if remote:
        first_remote = min(m.boundary[p]["startFace"] for p in remote)
        last_local = max([m.boundary[p]["startFace"] + m.boundary[p]["nFaces"] for p in local] + [nIF])
        assert first_remote >= last_local, "processor patches must follow all physical patches"
    m.localPatches, m.remotePatches = local, remote
    m.sortedPatches = sorted(local)
    m.nLocalCells = nLocalCells
    m.nRemoteCells = m.nCells - nLocalCells
    m.nLocalFaces = nLocalCells - nIC + nIF

    # --- deltas, weights, reconstruction weights (cmesh.cpp:168-233)
    Pc, Nc = cellCentres[owner], cellCentres[neighbour]
    delta = Pc - Nc
    d = np.sqrt(delta[:, 0] * delta[:, 0] + delta[:, 1] * delta[:, 1] + delta[:, 2] * delta[:, 2])
    deltasUnit = -delta / d[:, None]
    nFv = faceCentres - Nc
    pFv = faceCentres - Pc
    nD = np.abs((nFv[:, 0] * nrm[:, 0] + nFv[:, 1] * nrm[:, 1]) + nFv[:, 2] * nrm[:, 2])
    pD = np.abs((pFv[:, 0] * nrm[:, 0] + pFv[:, 1] * nrm[:, 1]) + pFv[:, 2] * nrm[:, 2])
    weights = nD / (nD + pD)
    w1 = ((-delta[:, 0] * pFv[:, 0]) + (-delta[:, 1] * pFv[:, 1])) + (-delta[:, 2] * pFv[:, 2])
    w2 = ((delta[:, 0] * nFv[:, 0]) + (delta[:, 1] * nFv[:, 1])) + (delta[:, 2] * nFv[:, 2])
    d2 = (delta[:, 0] * delta[:, 0] + delta[:, 1] * delta[:, 1]) + delta[:, 2] * delta[:, 2]
    w1 = w1 / d2
    w2 = w2 / d2
    linW = np.stack([w1 / 3, w2 / 3], axis=1)
    quadW = np.stack([2. / 3 * pFv + 1. / 3 * (pFv + w1[:, None] * delta),
                      2. / 3 * nFv + 1. / 3 * (nFv - w2[:, None] * delta)], axis=1)

    # --- cellOwner / cellNeighbours (cmesh.cpp:244-269); cellNeighbours = the "Full" variant
    fo = owner[cellFaces]
    fn = neighbour[cellFaces]
    own_flag = (fo == np.arange(nIC)[:, None])
    cellNeighbours = np.where(own_flag, fn, fo)

    m.points, m.faces = points, poly.faces
    m.owner = np.ascontiguousarray(owner, np.int32)
    m.neighbour = np.ascontiguousarray(neighbour, np.int32)
    m.cellFaces = np.ascontiguousarray(cellFaces, np.int32)
    m.cellNeighbours = np.ascontiguousarray(cellNeighbours, np.int32)
    m.cellOwner = np.ascontiguousarray(own_flag, np.int32)
    m.normals = np.ascontiguousarray(nrm)
    m.faceCentres = faceCentres
    m.cellCentres = cellCentres
    m.areas = area.reshape(-1, 1)
    m.volumes = vol.reshape(-1, 1)
    m.volumesL = m.volumes[m.owner]
    m.volumesR = m.volumes[m.neighbour[:nIF]]
    m.deltas = d.reshape(-1, 1)
    m.deltasUnit = np.ascontiguousarray(deltasUnit)
    m.weights = weights.reshape(-1, 1)
    m.linearWeights = np.ascontiguousarray(linW)
    m.quadraticWeights = np.ascontiguousarray(quadW)
    return m


def uniform_box(n, lo=(0., 0., 0.), hi=(1., 1., 1.), patches=None):
    """Closed-form MeshData of an undeformed nx*ny*nz box — the same arrays `build_mesh(hexmesh.box_mesh(...))`
    produces (checked to round-off in tests/test_metrics.py), without building points/faces or sorting: the
    benchmark sizes of BASELINE.json (up to 368^3 cells per GPU) are generated in seconds instead of minutes.
    patches: as hexmesh.box_mesh (default: six cyclic patches). Processor patches get the periodic-image / next-block
    cell centre as ghost centre (what the reference exchanges in createGhostCells, adFVM/mesh.py:784-805)."""
    from .hexmesh import SIDES
    nx, ny, nz = [int(v) for v in n]
    if patches is None:
        patches = [("x1", "cyclic", ["x-"], {"neighbourPatch": "x2"}), ("x2", "cyclic", ["x+"], {"neighbourPatch": "x1"}),
                   ("y1", "cyclic", ["y-"], {"neighbourPatch": "y2"}), ("y2", "cyclic", ["y+"], {"neighbourPatch": "y1"}),
                   ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}), ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})]
    lo = np.asarray(lo, np.float64); hi = np.asarray(hi, np.float64)
    h = (hi - lo) / np.array([nx, ny, nz], np.float64)
    C = nx * ny * nz
    stride = np.array([1, nx, nx * ny], np.int64)
    c = np.arange(C, dtype=np.int64)
    I, J, K = c % nx, (c // nx) % ny, c // (nx * ny)
    idx = (I, J, K); dims = (nx, ny, nz)
    has = [idx[d] < dims[d] - 1 for d in range(3)]
    cnt = has[0].astype(np.int64) + has[1] + has[2]
    start = np.cumsum(cnt) - cnt
    fid = [start, start + has[0], start + has[0] + has[1]]           # id of the cell's +x/+y/+z internal face
    Fi = int(cnt.sum())
    # boundary faces, patch by patch, side by side
    def side_cells(side):
        d = "xyz".index(side[0]); a, b = [q for q in range(3) if q != d]
        Bb, Aa = np.meshgrid(np.arange(dims[b]), np.arange(dims[a]), indexing="ij")
        ijk = [None, None, None]
        ijk[a], ijk[b] = Aa.ravel(), Bb.ravel()
        ijk[d] = np.zeros_like(ijk[a]) if side[1] == "-" else np.full_like(ijk[a], dims[d] - 1)
        return d, ijk[0] + nx * (ijk[1] + ny * ijk[2])
    boundary = OrderedDict()
    b_owner, b_dim, b_sign, b_coupled = [], [], [], []
    used = set(); startFace = Fi
    side_list = []
    for name, ptype, sides, extra in patches:
        n_p = 0
        for s in sides:
            assert s in SIDES and s not in used, s
            used.add(s)
            d, cells = side_cells(s)
            side_list.append((startFace + n_p, cells))
            b_owner.append(cells); b_dim.append(np.full(len(cells), d)); b_sign.append(np.full(len(cells), -1. if s[1] == "-" else 1.))
            b_coupled.append(np.full(len(cells), ptype in COUPLED_PATCHES))
            n_p += len(cells)
        dct = OrderedDict(type=ptype, nFaces=int(n_p), startFace=int(startFace)); dct.update(extra or {})
        boundary[name] = dct
        startFace += n_p
    assert used == set(SIDES)
    b_owner = np.concatenate(b_owner); b_dim = np.concatenate(b_dim); b_sign = np.concatenate(b_sign); b_coupled = np.concatenate(b_coupled)
    G = len(b_owner); F = Fi + G
    m = MeshData()
    m.boundary = boundary
    m.nFaces, m.nInternalFaces, m.nInternalCells, m.nGhostCells, m.nCells, m.nBoundaryFaces = F, Fi, C, G, C + G, G
    owner = np.empty(F, np.int64); neigh = np.empty(F, np.int64); fdim = np.empty(F, np.int8)
    # the internal faces are numbered cell by cell, +x, +y, +z: the (cell, direction) pairs of the existing ones, in row-major order
    cells_idx, dirs = np.nonzero(np.stack(has, axis=1))
    owner[:Fi] = cells_idx; neigh[:Fi] = cells_idx + stride[dirs]; fdim[:Fi] = dirs
    del cells_idx, dirs
    owner[Fi:] = b_owner; neigh[Fi:] = C + np.arange(G); fdim[Fi:] = b_dim
    sign = np.ones(F); sign[Fi:] = b_sign
    coupled = np.ones(F, bool); coupled[Fi:] = b_coupled
    # patches: local / remote split, cellStartFace
    delta_cf = C - Fi
    local, remote = [], []
    nLocalCells = C
    for pid, patch in boundary.items():
        patch["cellStartFace"] = patch["startFace"] + delta_cf
        (remote if patch["type"] in PROCESSOR_PATCHES else local).append(pid)
        if patch["type"] not in PROCESSOR_PATCHES:
            nLocalCells += patch["nFaces"]
    if remote:
        assert min(boundary[p]["startFace"] for p in remote) >= max([boundary[p]["startFace"] + boundary[p]["nFaces"] for p in local] + [Fi])
    m.localPatches, m.remotePatches, m.sortedPatches = local, remote, sorted(local)
    m.nLocalCells, m.nRemoteCells, m.nLocalFaces = nLocalCells, C + G - nLocalCells, nLocalCells - C + Fi
    # geometry
    hf = h[fdim]                                                  # spacing normal to the face
    area = (h[0] * h[1] * h[2]) / hf
    normals = np.eye(3)[fdim] * sign[:, None]
    cc = np.empty((C + G, 3))
    cc[:C, 0] = lo[0] + (I + 0.5) * h[0]; cc[:C, 1] = lo[1] + (J + 0.5) * h[1]; cc[:C, 2] = lo[2] + (K + 0.5) * h[2]
    fc = cc[owner]; fc += (0.5 * hf)[:, None] * normals
    gh = fc[Fi:].copy()
    bc = coupled[Fi:]
    gh[bc, b_dim[bc]] += 0.5 * hf[Fi:][bc] * b_sign[bc]           # coupled: the image cell's centre; else the face centre
    cc[C:] = gh
    dist = np.where(coupled, hf, 0.5 * hf)
    w1 = np.where(coupled, 0.5, 1.0); w2 = np.where(coupled, 0.5, 0.0)
    # F - P = (hf/2) n, F - N = -(hf/2) n on coupled faces (0 elsewhere), P - N = -dist n: every vector below is a scalar times n
    m.owner = np.ascontiguousarray(owner, np.int32); m.neighbour = np.ascontiguousarray(neigh, np.int32)
    m.normals = normals; m.faceCentres = fc; m.cellCentres = cc
    m.areas = area.reshape(-1, 1)
    m.volumes = np.full((C, 1), h[0] * h[1] * h[2])
    m.volumesL = m.volumes[m.owner]; m.volumesR = m.volumes[m.neighbour[:Fi]]
    m.deltas = dist.reshape(-1, 1); m.deltasUnit = normals.copy()
    m.weights = np.where(coupled, 0.5, 0.0).reshape(-1, 1)
    m.linearWeights = np.stack([w1 / 3, w2 / 3], axis=1)
    qw = np.empty((F, 2, 3))
    np.multiply((0.5 * hf - (w1 / 3) * dist)[:, None], normals, out=qw[:, 0, :])                 # (F-P) + (w1/3)(P-N)
    np.multiply((np.where(coupled, -0.5 * hf, 0.0) + (w2 / 3) * dist)[:, None], normals, out=qw[:, 1, :])   # (F-N) - (w2/3)(P-N)
    m.quadraticWeights = qw
    # cellFaces: owned faces ascending (internal +x,+y,+z, then boundary in face order), then neighbour-side faces ascending
    cellFaces = np.full((C, 6), -1, np.int64); fill = np.zeros(C, np.int64)
    def put(cells, faces):
        cellFaces[cells, fill[cells]] = faces; fill[cells] += 1
    # cells that touch no side of the box (all but O(n^2)): [+x, +y, +z, -z, -y, -x] by strided copies of the face-id arrays
    inner = np.zeros((nz, ny, nx), bool)
    if min(nx, ny, nz) > 2:
        inner[1:-1, 1:-1, 1:-1] = True
        cf4 = cellFaces.reshape(nz, ny, nx, 6)
        f3 = [fid[d].reshape(nz, ny, nx) for d in range(3)]
        for d in range(3):
            cf4[1:-1, 1:-1, 1:-1, d] = f3[d][1:-1, 1:-1, 1:-1]
        cf4[1:-1, 1:-1, 1:-1, 3] = f3[2][0:-2, 1:-1, 1:-1]
        cf4[1:-1, 1:-1, 1:-1, 4] = f3[1][1:-1, 0:-2, 1:-1]
        cf4[1:-1, 1:-1, 1:-1, 5] = f3[0][1:-1, 1:-1, 0:-2]
        fill[inner.ravel()] = 6
    rim = ~inner.ravel()
    for d in range(3):
        sel = has[d] & rim
        put(c[sel], fid[d][sel])
    for s0, cells in side_list:
        put(cells, s0 + np.arange(len(cells)))
    for d in (2, 1, 0):
        sel = (idx[d] > 0) & rim
        put(c[sel], fid[d][c[sel] - stride[d]])
    assert np.all(fill == 6)
    fo = owner[cellFaces]; fn = neigh[cellFaces]
    own_flag = fo == c[:, None]
    m.cellFaces = np.ascontiguousarray(cellFaces, np.int32)
    m.cellNeighbours = np.ascontiguousarray(np.where(own_flag, fn, fo), np.int32)
    m.cellOwner = np.ascontiguousarray(own_flag, np.int32)
    return m


def _build_mesh_device(poly, remote_centres, dev):
    """Connectivity and patch bookkeeping on the host (integer sorts), geometry by adfvm_mesh_metrics."""
    import ctypes as C
    from . import _lib as L
    lib = dev.get("lib") or L.default_lib()
    dtype = np.dtype(dev.get("precision", np.float64))
    m = MeshData()
    faces = np.ascontiguousarray(poly.faces, np.int32)
    owner = np.ascontiguousarray(poly.owner, np.int32)
    nb_int = np.ascontiguousarray(poly.neighbour, np.int32)
    nF, nIF = len(owner), len(nb_int)
    nIC = int(owner.max()) + 1
    nG = nF - nIF
    m.boundary = OrderedDict((k, dict(v)) for k, v in poly.boundary.items())
    m.nFaces, m.nInternalFaces, m.nInternalCells = nF, nIF, nIC
    m.nGhostCells, m.nCells, m.nBoundaryFaces = nG, nIC + nG, nG
    cells_of_face = np.concatenate([owner.astype(np.int64), nb_int.astype(np.int64)])
    face_ids = np.concatenate([np.arange(nF), np.arange(nIF)])
    pri = np.concatenate([np.zeros(nF, np.int64), np.ones(nIF, np.int64)])
    order = np.lexsort((face_ids, pri, cells_of_face))
    assert len(order) == 6 * nIC, "hexahedral cells only (6 faces per cell)"
    cellFaces = np.ascontiguousarray(face_ids[order].reshape(nIC, 6), np.int32)
    neighbour = np.concatenate([nb_int, np.arange(nIC, nIC + nG, dtype=np.int32)]).astype(np.int32)
    delta_cf = nIC - nIF
    pids = sorted(m.boundary.keys(), key=lambda x: (m.boundary[x]["startFace"], m.boundary[x]["nFaces"]))
    local, remote = [], []
    nLocalCells = nIC
    table = (L.MetricPatch * max(1, len(pids)))()
    rc = np.zeros((nG, 3), dtype) if any(m.boundary[p]["type"] in PROCESSOR_PATCHES for p in pids) else None
    for i, pid in enumerate(pids):
        patch = m.boundary[pid]
        patch["nFaces"] = int(patch["nFaces"]); patch["startFace"] = int(patch["startFace"])
        patch["cellStartFace"] = patch["startFace"] + delta_cf
        t = table[i]
        t.startFace, t.nFaces, t.kind, t.nbrStartFace = patch["startFace"], patch["nFaces"], 0, 0
        if patch["type"] in PROCESSOR_PATCHES:
            remote.append(pid); t.kind = 2
            assert remote_centres is not None and pid in remote_centres, "processor patch %s needs the neighbour rank's cell centres" % pid
            s0 = patch["startFace"] - nIF
            rc[s0:s0 + patch["nFaces"]] = remote_centres[pid]
        else:
            local.append(pid); nLocalCells += patch["nFaces"]
            if patch["type"] in CYCLIC_PATCHES:
                t.kind = 1; t.nbrStartFace = int(m.boundary[patch["neighbourPatch"]]["startFace"])
    m.localPatches, m.remotePatches, m.sortedPatches = local, remote, sorted(local)
    m.nLocalCells, m.nRemoteCells = nLocalCells, nIC + nG - nLocalCells
    m.nLocalFaces = nLocalCells - nIC + nIF
    points = np.ascontiguousarray(poly.points, dtype)
    out = {"areas": (nF, 1), "normals": (nF, 3), "faceCentres": (nF, 3), "cellCentres": (nIC + nG, 3), "volumes": (nIC, 1),
           "deltas": (nF, 1), "deltasUnit": (nF, 3), "weights": (nF, 1), "linearWeights": (nF, 2), "quadraticWeights": (nF, 2, 3)}
    arr = {k: np.zeros(shp, dtype) for k, shp in out.items()}
    ctx = C.c_void_p()
    stream = dev.get("stream")
    lib.check(lib.dll.adfvm_create(C.byref(ctx), int(dev.get("device", 0)), dtype.itemsize, C.c_void_p(stream) if stream else None))
    try:
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        lib.check(lib.dll.adfvm_mesh_metrics(ctx, len(points), p(points), nF, nIF, nIC, p(faces), p(owner), p(neighbour), p(cellFaces),
                                             len(pids), table, p(rc), *[p(arr[k]) for k in out]))
    finally:
        lib.dll.adfvm_destroy(ctx)
    for k, v in arr.items():
        setattr(m, k, v)
    own_flag = (owner[cellFaces] == np.arange(nIC)[:, None])
    m.points, m.faces = poly.points, poly.faces
    m.owner, m.neighbour, m.cellFaces = owner, neighbour, cellFaces
    m.cellNeighbours = np.ascontiguousarray(np.where(own_flag, neighbour[cellFaces], owner[cellFaces]), np.int32)
    m.cellOwner = np.ascontiguousarray(own_flag, np.int32)
    m.volumesL = m.volumes[m.owner]
    m.volumesR = m.volumes[m.neighbour[:nIF]]
    return m
