"""Multi-block structured hexahedral mesher in the conventions of OpenFOAM's blockMesh: the meshes of the reference's shipped
cases are given as `blockMeshDict`s (cases/cylinder, cases/forwardStep, cases/vane_optim/foam/laminar: vertices, hex blocks with
cell counts and simpleGrading, arc / spline edges, patches as lists of block faces) and OpenFOAM is not available where this
runs. `block_mesh` produces the polyMesh those dictionaries describe, in OpenFOAM's layout:

  * cells block by block, i fastest inside a block; coincident points of neighbouring blocks merged;
  * internal faces in upper-triangular order (by owner, then neighbour), normal from owner to neighbour;
  * boundary faces patch by patch in dictionary order, block faces not named by any patch collected in `defaultFaces`
    (type empty) - so `startFace` / `nFaces` / `nInternalFaces` can be checked against the `boundary` files the reference ships
    (tests/test_blockmesh.py: cases/cylinder and cases/forwardStep match exactly);
  * points from the twelve block edges (straight, circular arc through a mid point, spline through given points; geometric
    grading along every edge) by transfinite interpolation. OpenFOAM's own point formula differs in how it blends the edge
    parameters inside a block; topology, numbering and patches do not depend on it.

Block-local vertex order as in blockMesh: v0..v3 the z- face counter-clockwise seen from +z ((0,0),(1,0),(1,1),(0,1)), v4..v7
the z+ face in the same order. Output: hexmesh.PolyMesh (points, faces [F][4], owner, neighbour, boundary).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from .hexmesh import PolyMesh

# (a, b) block-local end vertices of the 12 edges: 0-3 along x (at (y,z) = (0,0),(1,0),(1,1),(0,1)), 4-7 along y
# (at (x,z) = (0,0),(1,0),(1,1),(0,1)), 8-11 along z (at (x,y) = (0,0),(1,0),(1,1),(0,1))
EDGES = [(0, 1), (3, 2), (7, 6), (4, 5), (0, 3), (1, 2), (5, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
# block faces: side -> the four block-local vertices
SIDES = {"x-": (0, 4, 7, 3), "x+": (1, 2, 6, 5), "y-": (0, 1, 5, 4), "y+": (3, 7, 6, 2), "z-": (0, 3, 2, 1), "z+": (4, 5, 6, 7)}


def _grading(n, ratio):
    """parameter values of n cells whose sizes grow geometrically, last/first = ratio (simpleGrading)"""
    if n == 1 or abs(ratio - 1.0) < 1e-12:
        return np.linspace(0., 1., n + 1)
    k = ratio ** (1.0 / (n - 1))
    return (1 - k ** np.arange(n + 1)) / (1 - k ** n)


def _arc(p0, p1, pm, s):
    """points at arc-length fractions s of the circle through p0, pm, p1"""
    a, b = pm - p0, p1 - p0
    nrm = np.cross(a, b)
    # circumcentre of the triangle (p0, pm, p1)
    c = p0 + np.cross(np.dot(a, a) * b - np.dot(b, b) * a, nrm) / (2 * np.dot(nrm, nrm))
    r0, r1 = p0 - c, p1 - c
    R = np.linalg.norm(r0)
    ang = np.arccos(np.clip(np.dot(r0, r1) / (R * np.linalg.norm(r1)), -1, 1))
    ax = nrm / np.linalg.norm(nrm)
    e1 = r0 / R
    e2 = np.cross(ax, e1)
    t = s * ang
    return c + R * (np.cos(t)[:, None] * e1 + np.sin(t)[:, None] * e2)


def _spline(pts, s):
    """points at arc-length fractions s of a cubic spline through pts (chord-length parametrisation)"""
    from scipy.interpolate import CubicSpline
    pts = np.asarray(pts, np.float64)
    t = np.concatenate([[0.], np.cumsum(np.linalg.norm(np.diff(pts, axis=0), axis=1))])
    cs = CubicSpline(t / t[-1], pts, bc_type="natural")
    # re-parametrise by arc length
    fine = np.linspace(0, 1, 2001)
    xyz = cs(fine)
    L = np.concatenate([[0.], np.cumsum(np.linalg.norm(np.diff(xyz, axis=0), axis=1))])
    return cs(np.interp(s * L[-1], L, fine))


def block_mesh(vertices, blocks, edges=(), patches=(), scale=1.0, default_patch=("defaultFaces", "empty")):
    """vertices: [nV][3]; blocks: [(v[8], (nx, ny, nz), (gx, gy, gz))]; edges: [("arc", a, b, mid) | ("spline", a, b, pts)];
    patches: [(name, type, [quad of global vertex ids, ...], extra dict)]"""
    V = np.asarray(vertices, np.float64) * scale
    curved = {}
    for e in edges:
        kind, a, b, data = e
        curved[(a, b)] = (kind, np.asarray(data, np.float64) * scale, False)
        curved[(b, a)] = (kind, np.asarray(data, np.float64) * scale, True)

    def edge_points(a, b, s):
        if (a, b) in curved:
            kind, data, rev = curved[(a, b)]
            p0, p1 = (V[b], V[a]) if rev else (V[a], V[b])
            ss = (1 - s)[::-1] if rev else s
            pts = _arc(p0, p1, data, ss) if kind == "arc" else _spline(np.vstack([p0, data, p1]) if not np.allclose(data[0], p0) else data, ss)
            return pts[::-1] if rev else pts
        return V[a] + s[:, None] * (V[b] - V[a])

    all_pts, cell_off, blk_info = [], [0], []
    for bv, (nx, ny, nz), (gx, gy, gz) in blocks:
        n = (nx, ny, nz)
        s = [_grading(nx, gx), _grading(ny, gy), _grading(nz, gz)]
        E = [edge_points(bv[a], bv[b], s[k // 4]) for k, (a, b) in enumerate(EDGES)]
        c = V[list(bv)]
        u, v, w = np.meshgrid(s[0], s[1], s[2], indexing="ij")
        u, v, w = u[..., None], v[..., None], w[..., None]
        I = np.arange(nx + 1)[:, None, None]; J = np.arange(ny + 1)[None, :, None]; K = np.arange(nz + 1)[None, None, :]
        Ex = [E[k][I] for k in range(4)]; Ey = [E[4 + k][J] for k in range(4)]; Ez = [E[8 + k][K] for k in range(4)]
        P = (1 - v) * (1 - w) * Ex[0] + v * (1 - w) * Ex[1] + v * w * Ex[2] + (1 - v) * w * Ex[3] \
            + (1 - u) * (1 - w) * Ey[0] + u * (1 - w) * Ey[1] + u * w * Ey[2] + (1 - u) * w * Ey[3] \
            + (1 - u) * (1 - v) * Ez[0] + u * (1 - v) * Ez[1] + u * v * Ez[2] + (1 - u) * v * Ez[3]
        tri = (1 - u) * (1 - v) * (1 - w) * c[0] + u * (1 - v) * (1 - w) * c[1] + u * v * (1 - w) * c[2] + (1 - u) * v * (1 - w) * c[3] \
            + (1 - u) * (1 - v) * w * c[4] + u * (1 - v) * w * c[5] + u * v * w * c[6] + (1 - u) * v * w * c[7]
        P = P - 2 * tri                                   # [nx+1][ny+1][nz+1][3]
        all_pts.append(P.reshape(-1, 3))
        blk_info.append((n, len(all_pts) - 1))
        cell_off.append(cell_off[-1] + nx * ny * nz)
    pts = np.vstack(all_pts)
    # ---- merge coincident points of neighbouring blocks
    span = pts.max(axis=0) - pts.min(axis=0)
    q = np.round((pts - pts.min(axis=0)) / (span.max() * 1e-9)).astype(np.int64)
    _, first, inv = np.unique(q, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)                              # keep the order of first appearance
    rank = np.empty_like(order); rank[order] = np.arange(len(order))
    gid = rank[inv.ravel()]
    points = pts[first[order]]
    # ---- cells and their six faces
    quads, fcell, fblock, fside = [], [], [], []
    poff = 0
    for b, (bv, (nx, ny, nz), _) in enumerate(blocks):
        def pid(i, j, k):
            return gid[poff + (i * (ny + 1) + j) * (nz + 1) + k]
        i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        cid = cell_off[b] + i + nx * (j + ny * k)
        c8 = [pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k),
              pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j + 1, k + 1), pid(i, j + 1, k + 1)]
        for side, lv in SIDES.items():
            quads.append(np.stack([c8[x].ravel() for x in lv], axis=1))
            fcell.append(cid.ravel())
            d = "xyz".index(side[0]); idx = (i, j, k)[d].ravel()
            on = (idx == 0) if side[1] == "-" else (idx == (nx, ny, nz)[d] - 1)
            fblock.append(np.where(on, b, -1)); fside.append(np.full(cid.size, list(SIDES).index(side)))
        poff += (nx + 1) * (ny + 1) * (nz + 1)
    quads = np.vstack(quads); fcell = np.concatenate(fcell); fblock = np.concatenate(fblock); fside = np.concatenate(fside)
    key = np.sort(quads, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    inv = inv.ravel()
    internal = cnt[inv] == 2
    # internal faces: owner = the lower cell; take the owner's (outward) quad; upper-triangular order
    o = np.argsort(inv[internal], kind="stable")
    ii = np.where(internal)[0][o].reshape(-1, 2)
    lo = np.where(fcell[ii[:, 0]] < fcell[ii[:, 1]], ii[:, 0], ii[:, 1]); hi = np.where(fcell[ii[:, 0]] < fcell[ii[:, 1]], ii[:, 1], ii[:, 0])
    own_i, nb_i = fcell[lo], fcell[hi]
    srt = np.lexsort((nb_i, own_i))
    faces_int, own_i, nb_i = quads[lo][srt], own_i[srt], nb_i[srt]
    # boundary faces by patch
    bidx = np.where(~internal)[0]
    assigned = np.zeros(len(bidx), bool)
    side_names = list(SIDES)
    boundary = OrderedDict()
    faces_b, own_b = [], []
    start = len(faces_int)

    def block_side(quad):
        qs = set(int(x) for x in quad)
        for b, (bv, _, _) in enumerate(blocks):
            for sname, lv in SIDES.items():
                if set(int(bv[x]) for x in lv) == qs:
                    return b, side_names.index(sname)
        raise ValueError("patch face %s is not a face of any block" % (quad,))
    for name, ptype, qlist, extra in patches:
        sel_all = []
        for quad in qlist:
            b, sd = block_side(quad)
            sel = np.where((fblock[bidx] == b) & (fside[bidx] == sd) & ~assigned)[0]
            assigned[sel] = True
            sel_all.append(sel)
        sel_all = np.concatenate(sel_all) if sel_all else np.zeros(0, np.int64)
        d = OrderedDict(type=ptype, nFaces=int(len(sel_all)), startFace=int(start)); d.update(extra or {})
        boundary[name] = d
        faces_b.append(quads[bidx[sel_all]]); own_b.append(fcell[bidx[sel_all]])
        start += len(sel_all)
    rest = np.where(~assigned)[0]
    if len(rest):
        boundary[default_patch[0]] = OrderedDict(type=default_patch[1], nFaces=int(len(rest)), startFace=int(start))
        faces_b.append(quads[bidx[rest]]); own_b.append(fcell[bidx[rest]])
    faces = np.vstack([faces_int] + faces_b)
    owner = np.concatenate([own_i] + own_b)
    # orientation: right-hand normal out of the owner
    P = points[faces]
    nrm = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 1])
    ncell = cell_off[-1]
    csum = np.zeros((ncell, 3)); ccnt = np.zeros(ncell)
    np.add.at(csum, fcell, points[quads].mean(axis=1)); np.add.at(ccnt, fcell, 1)
    cc = csum / ccnt[:, None]
    flip = ((P.mean(axis=1) - cc[owner]) * nrm).sum(axis=1) < 0
    faces[flip] = faces[flip][:, ::-1]
    return PolyMesh(points, faces.astype(np.int32), owner.astype(np.int32), nb_i.astype(np.int32), boundary)


# ------------------------------------------------------------------------------------------ the reference's shipped dictionaries
def forward_step_dict():
    """cases/forwardStep/constant/polyMesh/blockMeshDict: Mach-3 forward-facing step, 3 blocks, 16 128 cells"""
    xy = [(0, 0), (0.6, 0), (0, 0.2), (0.6, 0.2), (3, 0.2), (0, 1), (0.6, 1), (3, 1)]
    verts = [(x, y, -0.05) for x, y in xy] + [(x, y, 0.05) for x, y in xy]
    blocks = [((0, 1, 3, 2, 8, 9, 11, 10), (48, 16, 1), (1, 1, 1)), ((2, 3, 6, 5, 10, 11, 14, 13), (48, 64, 1), (1, 1, 1)),
              ((3, 4, 7, 6, 11, 12, 15, 14), (192, 64, 1), (1, 1, 1))]
    patches = [("inlet", "patch", [(0, 8, 10, 2), (2, 10, 13, 5)], {}), ("outlet", "patch", [(4, 7, 15, 12)], {}),
               ("bottom", "symmetryPlane", [(0, 1, 9, 8)], {}), ("top", "symmetryPlane", [(5, 13, 14, 6), (6, 14, 15, 7)], {}),
               ("obstacle", "patch", [(1, 3, 11, 9), (3, 4, 12, 11)], {})]
    return dict(vertices=verts, blocks=blocks, edges=[], patches=patches, scale=1.0)


def cylinder_dict(cyclic_span=False):
    """cases/cylinder/constant/polyMesh/blockMeshDict: upper half of the laminar cylinder (radius 0.5 x 2.5e-4 m), 10 blocks with
    arc edges, 46 250 cells. cyclic_span: name the two z planes z1 / z2 and make them cyclic, as the case's create_mesh.sh does
    with createPatch (the shipped `boundary` file still has them as patches z1noc / z2noc)."""
    xy = [(0.5, 0), (1, 0), (10, 0), (10, 0.707107), (0.707107, 0.707107), (0.353553, 0.353553), (10, 5), (0.707107, 5), (0, 5),
          (0, 1), (0, 0.5), (-0.5, 0), (-1, 0), (-3, 0), (-3, 0.707107), (-0.707107, 0.707107), (-0.353553, 0.353553), (-3, 5),
          (-0.707107, 5)]
    verts = [(x, y, -0.5) for x, y in xy] + [(x, y, 0.5) for x, y in xy]
    blocks = [((5, 4, 9, 10, 24, 23, 28, 29), (25, 25, 1), (10, 1, 1)), ((0, 1, 4, 5, 19, 20, 23, 24), (25, 25, 1), (10, 1, 1)),
              ((1, 2, 3, 4, 20, 21, 22, 23), (200, 25, 1), (1, 1, 1)), ((4, 3, 6, 7, 23, 22, 25, 26), (200, 125, 1), (1, 0.09, 1)),
              ((9, 4, 7, 8, 28, 23, 26, 27), (25, 125, 1), (1, 0.09, 1)), ((15, 16, 10, 9, 34, 35, 29, 28), (25, 25, 1), (0.1, 1, 1)),
              ((12, 11, 16, 15, 31, 30, 35, 34), (25, 25, 1), (0.1, 1, 1)), ((13, 12, 15, 14, 32, 31, 34, 33), (50, 25, 1), (1, 1, 1)),
              ((14, 15, 18, 17, 33, 34, 37, 36), (50, 125, 1), (1, 0.09, 1)), ((15, 9, 8, 18, 34, 28, 27, 37), (25, 125, 1), (1, 0.09, 1))]
    arcs = [(0, 5, (0.469846, 0.17101)), (5, 10, (0.17101, 0.469846)), (1, 4, (0.939693, 0.34202)), (4, 9, (0.34202, 0.939693)),
            (11, 16, (-0.469846, 0.17101)), (16, 10, (-0.17101, 0.469846)), (12, 15, (-0.939693, 0.34202)), (15, 9, (-0.34202, 0.939693))]
    edges = [("arc", a, b, (x, y, -0.5)) for a, b, (x, y) in arcs] + [("arc", a + 19, b + 19, (x, y, 0.5)) for a, b, (x, y) in arcs]
    z1 = [(0, 5, 4, 1), (5, 10, 9, 4), (10, 16, 15, 9), (16, 11, 12, 15), (12, 13, 14, 15), (15, 14, 17, 18), (15, 18, 8, 9), (9, 8, 7, 4),
          (4, 7, 6, 3), (1, 4, 3, 2)]
    z2 = [(19, 20, 23, 24), (24, 23, 28, 29), (29, 28, 34, 35), (35, 34, 31, 30), (31, 34, 33, 32), (34, 37, 36, 33), (34, 28, 27, 37),
          (28, 23, 26, 27), (23, 22, 25, 26), (20, 21, 22, 23)]
    zn = ("z1", "z2", "cyclic") if cyclic_span else ("z1noc", "z2noc", "patch")
    patches = [("down", "patch", [(0, 1, 20, 19), (1, 2, 21, 20), (12, 11, 30, 31), (13, 12, 31, 32)], {}),
               ("right", "patch", [(2, 3, 22, 21), (3, 6, 25, 22)], {}),
               ("up", "patch", [(7, 8, 27, 26), (6, 7, 26, 25), (8, 18, 37, 27), (18, 17, 36, 37)], {}),
               ("left", "patch", [(14, 13, 32, 33), (17, 14, 33, 36)], {}),
               ("cylinder", "patch", [(10, 5, 24, 29), (5, 0, 19, 24), (16, 10, 29, 35), (11, 16, 35, 30)], {}),
               (zn[0], zn[2], z1, {"neighbourPatch": zn[1]} if cyclic_span else {}),
               (zn[1], zn[2], z2, {"neighbourPatch": zn[0]} if cyclic_span else {})]
    return dict(vertices=verts, blocks=blocks, edges=edges, patches=patches, scale=2.5e-4)


def match_cyclic(poly, a, b, separation):
    """reorder the faces of cyclic patch `b` so that its face i is the periodic image of face i of patch `a`
    (centre_b = centre_a + separation): the pairing convention of the reference (adFVM/BCs.py:88-95)"""
    A, B = poly.boundary[a], poly.boundary[b]
    assert A["nFaces"] == B["nFaces"]
    n, sa, sb = A["nFaces"], A["startFace"], B["startFace"]
    ca = poly.points[poly.faces[sa:sa + n]].mean(axis=1) + np.asarray(separation, np.float64)
    cb = poly.points[poly.faces[sb:sb + n]].mean(axis=1)
    from scipy.spatial import cKDTree
    dist, idx = cKDTree(cb).query(ca)
    scale = np.abs(poly.points).max()
    assert dist.max() < 1e-6 * scale and len(np.unique(idx)) == n, "cyclic patches %s / %s do not match" % (a, b)
    poly.faces[sb:sb + n] = poly.faces[sb + idx]
    poly.owner[sb:sb + n] = poly.owner[sb + idx]
    return poly


def vane_dict(nz=1, span=1.0):
    """cases/vane_optim/foam/laminar/constant/polyMesh/blockMeshDict: one passage of the turbine-vane cascade (pitch 57.5 mm), 16
    blocks of 25 x 25 cells with spline edges along the two blade surfaces (10 000 cells per layer). The dictionary has an empty
    `boundary` list (the reference's patches were created afterwards); here: inlet, outlet, the blade surfaces `pressure` (blade
    above the passage) and `suction` (blade below), the pitchwise periodic pair midplane1 / midplane2 and the spanwise pair
    z1plane / z2plane, with the names and types of the reference's vane cases (cases/vane_optim/foam/laminar/constant/polyMesh/
    boundary). nz layers over `span` mm in z (the dictionary is one layer of 1 mm)."""
    xy = [(0.0, -57.49995), (3.895, -49.97595), (18.734, -52.79495), (26.338, -68.48795), (36.461, -109.80495), (0.0, 0.0), (4.081, -4.895),
          (28.935, -34.724), (35.427, -51.465), (36.461, -52.305), (-10.0, -57.49995), (0.895, -42.97595), (23.734, -47.79495),
          (33.338, -65.48795), (42.1967643635, -117.996470443), (-10.0, 0.0), (-2.919, -7.895), (26.935, -39.724), (32.427, -58.465),
          (42.1967643635, -60.4965204429), (-100.0, -57.49995), (-100.0, -45.99996), (-100.0, -11.49999), (-100.0, 0.0),
          (93.8186436351, -191.720154429), (93.8186436351, -180.220164429), (93.8186436351, -145.720194429), (93.8186436351, -134.220204429)]
    nv = len(xy)
    verts = [(x, y, 0.0) for x, y in xy] + [(x, y, -span) for x, y in xy]
    quads2d = [((0, 10, 11, 1), (10, 1)), ((1, 11, 12, 2), (10, 1)), ((2, 12, 13, 3), (10, 1)), ((3, 13, 14, 4), (10, 1)),
               ((5, 6, 16, 15), (1, 10)), ((6, 7, 17, 16), (1, 10)), ((7, 8, 18, 17), (1, 10)), ((8, 9, 19, 18), (1, 10)),
               ((22, 16, 11, 21), (1, 1)), ((16, 17, 12, 11), (1, 1)), ((17, 18, 13, 12), (1, 1)), ((18, 26, 25, 13), (1, 1)),
               ((21, 11, 10, 20), (1, 1)), ((23, 15, 16, 22), (1, 1)), ((19, 27, 26, 18), (1, 1)), ((13, 25, 24, 14), (1, 1))]
    blocks = [(tuple(q) + tuple(v + nv for v in q), (25, 25, nz), (g[0], g[1], 1)) for q, g in quads2d]
    splines = {(5, 6): [(0.0, 0.0), (0.185, -0.913), (0.927, -2.278), (1.669, -2.963), (2.782, -3.852), (4.081, -4.895)],
               (6, 7): [(4.081, -4.895), (9.089, -9.023), (13.64, -13.062), (18.363, -18.038), (22.444, -23.229), (25.967, -28.85), (28.935, -34.724)],
               (7, 8): [(28.935, -34.724), (31.347, -40.456), (33.757, -46.92), (34.87, -49.95), (35.427, -51.465)],
               (8, 9): [(35.427, -51.465), (36.075, -52.312), (36.461, -52.305)],
               (1, 0): [(3.895, -49.97595), (2.411, -51.64095), (1.484, -52.95195), (0.742, -54.20195), (0.185, -55.94595), (0.0, -57.49995)],
               (2, 1): [(18.734, -52.79495), (11.871, -47.71695), (3.895, -49.97595)],
               (3, 2): [(26.338, -68.48795), (23.0, -60.16295), (18.734, -52.79495)],
               (4, 3): [(36.461, -109.80495), (36.814, -109.58695), (36.726, -107.97595), (36.355, -106.44095), (35.241, -101.83695),
                        (33.201, -93.39695), (31.161, -85.11495), (28.75, -76.21295), (26.338, -68.48795)]}
    edges = []
    for (a, b), pts in splines.items():
        edges.append(("spline", a, b, [(x, y, 0.0) for x, y in pts]))
        edges.append(("spline", a + nv, b + nv, [(x, y, -span) for x, y in pts]))

    def wall(pairs):
        return [(a, b, b + nv, a + nv) for a, b in pairs]
    patches = [("midplane1", "cyclic", wall([(23, 15), (15, 5), (9, 19), (19, 27)]), {"neighbourPatch": "midplane2"}),
               ("midplane2", "cyclic", wall([(20, 10), (10, 0), (4, 14), (14, 24)]), {"neighbourPatch": "midplane1"}),
               ("z1plane", "cyclic", [q for q, _ in quads2d], {"neighbourPatch": "z2plane"}),
               ("z2plane", "cyclic", [tuple(v + nv for v in q) for q, _ in quads2d], {"neighbourPatch": "z1plane"}),
               ("inlet", "patch", wall([(20, 21), (21, 22), (22, 23)]), {}),
               ("outlet", "patch", wall([(24, 25), (25, 26), (26, 27)]), {}),
               ("pressure", "patch", wall([(5, 6), (6, 7), (7, 8), (8, 9)]), {}),
               ("suction", "patch", wall([(0, 1), (1, 2), (2, 3), (3, 4)]), {})]
    return dict(vertices=verts, blocks=blocks, edges=edges, patches=patches, scale=1e-3)


def start_faces_along(poly, normal):
    """rotate every face's vertex cycle (orientation kept) so that its first edge v0 -> v1 is the one most nearly parallel to
    the planes with the given normal. The reference's intersectPlane (adFVM/compat/cfuncs.pyx:60-69, see adfvm_b200.planecut)
    cuts a face split 2/2 through its edges (v0,v3) and (v2,v1) whatever the geometry, which is the right pair exactly when
    v0 -> v1 does not cross the plane; with this vertex order its cut areas are the geometric ones."""
    P = poly.points[poly.faces]                                       # [F][4][3]
    n = np.asarray(normal, np.float64)
    e01 = np.abs((P[:, 1] - P[:, 0]) @ n)
    e12 = np.abs((P[:, 2] - P[:, 1]) @ n)
    rot = e12 < e01
    poly.faces[rot] = np.roll(poly.faces[rot], -1, axis=1)
    return poly


def vane_mesh(nz=1, span=1.0, cut_normal=(1., 0., 0.)):
    poly = block_mesh(**vane_dict(nz, span))
    poly = match_cyclic(poly, "midplane1", "midplane2", (0., -57.49995e-3, 0.))
    return start_faces_along(poly, cut_normal) if cut_normal is not None else poly
