"""Structured hexahedral box meshes in adFVM's (OpenFOAM polyMesh) conventions.

The reference ships no `points/faces/owner/neighbour` files (they come from OpenFOAM's
`blockMesh`, reference `tests/setup_tests.sh:7-8`), so the synthetic benchmark meshes of
BASELINE.json and the parity fixtures are generated here, in memory, in exactly the layout
`Mesh.readFoam` produces (reference `adFVM/mesh.py:177-204`):

* internal faces first, in OpenFOAM upper-triangular order (sorted by owner, then neighbour),
* boundary patches contiguous after them, `startFace/nFaces` per patch,
* every face a quad whose first three points give a right-hand normal out of the owner
  (reference `adFVM/cpp/cmesh.cpp:72-87`),
* face *i* of a cyclic patch pairs with face *i* of its `neighbourPatch`
  (reference `adFVM/BCs.py:86-95`).

Cell id = i + nx*(j + ny*k); point id = i + (nx+1)*(j + (ny+1)*k).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

SIDES = ("x-", "x+", "y-", "y+", "z-", "z+")


class PolyMesh:
    """Plain container: points [nP,3] f64, faces [nF,4] i32, owner [nF] i32,
    neighbour [nInternalFaces] i32, boundary: OrderedDict name -> dict(type,nFaces,startFace,...)."""

    def __init__(self, points, faces, owner, neighbour, boundary):
        self.points = np.ascontiguousarray(points, np.float64)
        self.faces = np.ascontiguousarray(faces, np.int32)
        self.owner = np.ascontiguousarray(owner, np.int32)
        self.neighbour = np.ascontiguousarray(neighbour, np.int32)
        self.boundary = boundary

    @property
    def nInternalFaces(self):
        return len(self.neighbour)

    @property
    def nFaces(self):
        return len(self.owner)

    @property
    def nInternalCells(self):
        return int(self.owner.max()) + 1


def _axis_points(n, lo, hi, grading=None):
    if grading is None:
        return lo + (hi - lo) * np.arange(n + 1, dtype=np.float64) / n
    # geometric grading: ratio last/first cell size = grading
    r = grading ** (1.0 / max(n - 1, 1))
    w = r ** np.arange(n, dtype=np.float64)
    x = np.concatenate(([0.0], np.cumsum(w)))
    return lo + (hi - lo) * x / x[-1]


def box_mesh(n, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), patches=None, grading=None, warp=None):
    """Build an nx*ny*nz box.

    patches: list of (name, type, [sides], extra_dict). Sides not mentioned raise.
             Default: six cyclic patches x1/x2, y1/y2, z1/z2 (the §8(d) synthetic box).
    grading: optional (gx, gy, gz) geometric cell-size ratios (last/first).
    warp:    optional callable points[nP,3] -> points[nP,3] (smooth deformation; keeps topology),
             used to get non-orthogonal meshes that exercise every metric term.
    """
    nx, ny, nz = [int(v) for v in n]
    if patches is None:
        patches = [
            ("x1", "cyclic", ["x-"], {"neighbourPatch": "x2"}),
            ("x2", "cyclic", ["x+"], {"neighbourPatch": "x1"}),
            ("y1", "cyclic", ["y-"], {"neighbourPatch": "y2"}),
            ("y2", "cyclic", ["y+"], {"neighbourPatch": "y1"}),
            ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}),
            ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"}),
        ]
    g = grading or (None, None, None)
    xs = _axis_points(nx, lo[0], hi[0], g[0])
    ys = _axis_points(ny, lo[1], hi[1], g[1])
    zs = _axis_points(nz, lo[2], hi[2], g[2])
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    points = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if warp is not None:
        points = np.asarray(warp(points), np.float64)

    def pid(i, j, k):
        return (i + (nx + 1) * (j + (ny + 1) * k)).astype(np.int64)

    def cid(i, j, k):
        return (i + nx * (j + ny * k)).astype(np.int64)

    # quads with +x / +y / +z normals at point-plane index (i,j,k) = lower corner of the quad
    def xface(i, j, k):
        return np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], -1)

    def yface(i, j, k):
        return np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)], -1)

    def zface(i, j, k):
        return np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], -1)

    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    c = cid(I, J, K)
    # internal faces: per cell, to (i+1), (j+1), (k+1) neighbours (ascending neighbour id)
    cand_owner, cand_nb, cand_faces, cand_rank = [], [], [], []
    m = I < nx - 1
    cand_owner.append(c[m]); cand_nb.append(cid(I[m] + 1, J[m], K[m])); cand_faces.append(xface(I[m] + 1, J[m], K[m]))
    m = J < ny - 1
    cand_owner.append(c[m]); cand_nb.append(cid(I[m], J[m] + 1, K[m])); cand_faces.append(yface(I[m], J[m] + 1, K[m]))
    m = K < nz - 1
    cand_owner.append(c[m]); cand_nb.append(cid(I[m], J[m], K[m] + 1)); cand_faces.append(zface(I[m], J[m], K[m] + 1))
    own = np.concatenate(cand_owner)
    nb = np.concatenate(cand_nb)
    fcs = np.concatenate(cand_faces)
    order = np.lexsort((nb, own))
    own, nb, fcs = own[order], nb[order], fcs[order]

    # boundary sides; ordering inside a side: lexicographic in the two in-plane indices (fast index first)
    def side_faces(side):
        if side[0] == "x":
            Kk, Jj = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
            j, k = Jj.ravel(), Kk.ravel()
            if side[1] == "-":
                i = np.zeros_like(j)
                return cid(i, j, k), xface(i, j, k)[:, ::-1]
            i = np.full_like(j, nx - 1)
            return cid(i, j, k), xface(i + 1, j, k)
        if side[0] == "y":
            Kk, Ii = np.meshgrid(np.arange(nz), np.arange(nx), indexing="ij")
            i, k = Ii.ravel(), Kk.ravel()
            if side[1] == "-":
                j = np.zeros_like(i)
                return cid(i, j, k), yface(i, j, k)[:, ::-1]
            j = np.full_like(i, ny - 1)
            return cid(i, j, k), yface(i, j + 1, k)
        Jj, Ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        i, j = Ii.ravel(), Jj.ravel()
        if side[1] == "-":
            k = np.zeros_like(i)
            return cid(i, j, k), zface(i, j, k)[:, ::-1]
        k = np.full_like(i, nz - 1)
        return cid(i, j, k), zface(i, j, k + 1)

    used = set()
    boundary = OrderedDict()
    b_owner, b_faces = [], []
    start = len(own)
    for name, ptype, sides, extra in patches:
        o_list, f_list = [], []
        for s in sides:
            assert s in SIDES and s not in used, s
            used.add(s)
            o, f = side_faces(s)
            o_list.append(o); f_list.append(f)
        o = np.concatenate(o_list) if o_list else np.zeros(0, np.int64)
        f = np.concatenate(f_list) if f_list else np.zeros((0, 4), np.int64)
        d = OrderedDict(type=ptype, nFaces=int(len(o)), startFace=int(start))
        d.update(extra or {})
        boundary[name] = d
        b_owner.append(o); b_faces.append(f)
        start += len(o)
    assert used == set(SIDES), "every box side needs a patch: missing %s" % (set(SIDES) - used)
    owner = np.concatenate([own] + b_owner)
    faces = np.concatenate([fcs] + b_faces)
    return PolyMesh(points, faces, owner, nb, boundary)


def sine_warp(amp=0.03, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0)):
    """Smooth periodic-compatible warp that keeps the six bounding planes planar and periodic."""
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)

    def f(p):
        s = (p - lo) / (hi - lo)
        q = s.copy()
        tp = 2 * np.pi
        q[:, 0] += amp * np.sin(tp * s[:, 0]) * np.sin(tp * s[:, 1]) * np.sin(tp * s[:, 2])
        q[:, 1] += amp * np.sin(tp * s[:, 0]) * np.sin(2 * tp * s[:, 1]) * np.sin(tp * s[:, 2])
        q[:, 2] += 0.5 * amp * np.sin(2 * tp * s[:, 0]) * np.sin(tp * s[:, 1]) * np.sin(tp * s[:, 2])
        return lo + q * (hi - lo)

    return f


def masked_box_mesh(n, lo, hi, keep, patches, hole=("obstacle", "patch", {}), grading=None):
    """Structured nx*ny*nz box with the cells where `keep[k, j, i]` is False removed (e.g. the forward-facing step
    of reference cases/forwardStep/constant/polyMesh/blockMeshDict, whose three blocks are a box minus the step).
    patches: as in box_mesh, for the faces on the six box sides; faces between a kept and a removed cell go to the
    patch `hole` = (name, type, extra), placed after them. Unused points are kept (harmless to the readers)."""
    nx, ny, nz = [int(v) for v in n]
    keep = np.asarray(keep, bool).reshape(nz, ny, nx)
    g = grading or (None, None, None)
    xs = _axis_points(nx, lo[0], hi[0], g[0]); ys = _axis_points(ny, lo[1], hi[1], g[1]); zs = _axis_points(nz, lo[2], hi[2], g[2])
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    points = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def pid(i, j, k):
        return (i + (nx + 1) * (j + (ny + 1) * k)).astype(np.int64)

    quad = {0: lambda i, j, k: np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], -1),
            1: lambda i, j, k: np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)], -1),
            2: lambda i, j, k: np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], -1)}
    cmap = np.full((nz + 2, ny + 2, nx + 2), -1, np.int64)         # padded: index + 1
    cmap[1:-1, 1:-1, 1:-1][keep] = np.arange(int(keep.sum()))
    inside = np.zeros((nz + 2, ny + 2, nx + 2), bool); inside[1:-1, 1:-1, 1:-1] = True
    int_own, int_nb, int_f = [], [], []
    side_own = {s: [] for s in SIDES}; side_f = {s: [] for s in SIDES}
    hole_own, hole_f = [], []
    for d, ax in ((0, 2), (1, 1), (2, 0)):                          # direction d runs along padded-array axis ax
        nd = (nx, ny, nz)[d]
        slA = [slice(1, -1)] * 3; slB = [slice(1, -1)] * 3
        slA[ax] = slice(0, nd + 1); slB[ax] = slice(1, nd + 2)        # A = cell before the plane, B = cell after it
        A, B = cmap[tuple(slA)], cmap[tuple(slB)]
        inA, inB = inside[tuple(slA)], inside[tuple(slB)]
        shp = A.shape
        K, J, I = np.meshgrid(np.arange(shp[0]), np.arange(shp[1]), np.arange(shp[2]), indexing="ij")   # plane / cell indices
        A, B, inA, inB, I, J, K = [a.ravel() for a in (A, B, inA, inB, I, J, K)]
        f = quad[d](I, J, K)                                         # the index along d is already the plane index
        both = (A >= 0) & (B >= 0)
        int_own.append(A[both]); int_nb.append(B[both]); int_f.append(f[both])
        onlyA = (A >= 0) & (B < 0); onlyB = (B >= 0) & (A < 0)
        name = "xyz"[d]
        m = onlyA & ~inB; side_own[name + "+"].append(A[m]); side_f[name + "+"].append(f[m])
        m = onlyB & ~inA; side_own[name + "-"].append(B[m]); side_f[name + "-"].append(f[m][:, ::-1])
        m = onlyA & inB; hole_own.append(A[m]); hole_f.append(f[m])
        m = onlyB & inA; hole_own.append(B[m]); hole_f.append(f[m][:, ::-1])
    own = np.concatenate(int_own); nb = np.concatenate(int_nb); fcs = np.concatenate(int_f)
    order = np.lexsort((nb, own))
    own, nb, fcs = own[order], nb[order], fcs[order]
    boundary = OrderedDict()
    b_owner, b_faces = [], []
    start = len(own)
    used = set()
    plist = list(patches) + [(hole[0], hole[1], None, hole[2])]
    for name, ptype, sides, extra in plist:
        if sides is None:
            o = np.concatenate(hole_own); f = np.concatenate(hole_f)
        else:
            for s in sides:
                assert s in SIDES and s not in used, s
                used.add(s)
            o = np.concatenate([x for s in sides for x in side_own[s]] or [np.zeros(0, np.int64)])
            f = np.concatenate([x for s in sides for x in side_f[s]] or [np.zeros((0, 4), np.int64)])
        dct = OrderedDict(type=ptype, nFaces=int(len(o)), startFace=int(start))
        dct.update(extra or {})
        boundary[name] = dct
        b_owner.append(o); b_faces.append(f.reshape(-1, 4))
        start += len(o)
    assert used == set(SIDES), "every box side needs a patch: missing %s" % (set(SIDES) - used)
    return PolyMesh(points, np.concatenate([fcs] + b_faces), np.concatenate([own] + b_owner), nb, boundary)
