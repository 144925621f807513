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


def build_mesh(poly, remote_centres=None):
    """poly: hexmesh.PolyMesh (rank-local). remote_centres: for processor patches, dict
    patchID -> [nFaces,3] ghost cell centres (what the reference exchanges over MPI in
    `createGhostCells`, mesh.py:784-805)."""
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
