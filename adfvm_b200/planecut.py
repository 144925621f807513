"""Cells cut by a plane and the area of the cut through each: what the reference's vane design objective integrates over
(adFVM/objectives/vane.py:25-38 `getPlane` -> adFVM/compat/cfuncs.pyx:20-113 `intersectPlane`, a Cython function), restated
with numpy. The result is passed to the step functions as extraArgs (nPlaneCells, cells [n][1] int32, areas [n][1]).

Algorithm of the reference (quadrilateral faces):
  * a face is cut when some, but not all, of its four points lie on the positive side of the plane; the cut goes through two
    of its edges - for 1 / 3 points on one side the two edges at the odd point, for 2 / 2 the two edges that change side;
  * the two intersection points of every cut face; every cell collects the segments of its cut faces (3 or 4 of them);
  * area of the cell's cut polygon = half the norm of the cross product of its two "diagonals", built from the end points of
    the first segment and the segment end points nearest to them (cfuncs.pyx:101-111), as in the reference.
Cells are returned in ascending order (np.unique), like the reference.

Faithful to the reference's ACTUAL behaviour, which the parity tests pin against its compiled function: `left` holds the signed
distances (floats), not booleans, so the tests `left[j,k] == truth` (cfuncs.pyx:56) and `left[j,(k-1)%d] != left[j,k]` (:66)
compare distances - the odd-point search practically never matches and falls through to k = d-1, and the 2/2 case always takes
the edges (0,3) and (2,1). On the extruded structured meshes of the vane cases with a plane x = const this picks the cut edges
for the faces that matter; the quirk is reproduced, not corrected, because the objective value depends on it."""
from __future__ import annotations

import numpy as np


def intersect_plane(mesh, point, normal):
    """mesh: object with points [nP][3], faces [nF][4] (point ids; a leading count column as in the reference's [nF][5] is
    accepted), owner, neighbour, nInternalCells, nInternalFaces. Returns (cells int32 [n], areas [n])."""
    points = np.asarray(mesh.points, np.float64)
    faces = np.asarray(mesh.faces)
    if faces.shape[1] == 5:
        faces = faces[:, 1:]
    owner, neighbour = np.asarray(mesh.owner), np.asarray(mesh.neighbour)
    nIF = int(mesh.nInternalFaces)
    point, normal = np.asarray(point, np.float64), np.asarray(normal, np.float64)
    d = 4
    dist = (points[faces] - point) @ normal                          # [nF][4] signed distances (the reference's `left`)
    counter = (dist > 0.).sum(axis=1)
    inter = np.where((counter > 0) & (counter < d))[0]
    n = len(inter)
    L = dist[inter]
    lines = -np.ones((n, 4), np.int64)
    odd = (counter[inter] == 1) | (counter[inter] == d - 1)
    truth = (counter[inter] == 1).astype(np.float64)
    # odd point: the first k with left[j,k] == truth (a distance compared with 0. / 1.), else the loop ends at k = d-1
    hit = L == truth[:, None]
    k_odd = np.where(hit.any(axis=1), np.argmax(hit, axis=1), d - 1)
    lines[odd, 0] = k_odd[odd]; lines[odd, 1] = (k_odd[odd] - 1) % d
    lines[odd, 2] = k_odd[odd]; lines[odd, 3] = (k_odd[odd] + 1) % d
    ev = ~odd
    lines[ev, 0] = 0; lines[ev, 2] = 2
    for k in (0, 2):
        differs = L[:, (k - 1) % d] != L[:, k]
        lines[ev, k + 1] = np.where(differs[ev], (k - 1) % d, (k + 1) % d)
    ip = np.zeros((n, 2, 3))
    for i in (0, 2):
        l0 = points[faces[inter, lines[:, i]]]
        l1 = points[faces[inter, lines[:, i + 1]]]
        l = l1 - l0
        t = (((point - l0) @ normal) / (l @ normal)).reshape(-1, 1)
        ip[:, i // 2, :] = l0 + t * l
    internal = inter < nIF
    cells = np.unique(np.concatenate([owner[inter], neighbour[inter[internal]]]))
    cmap = np.zeros(int(mesh.nInternalCells), np.int64)
    cmap[cells] = np.arange(len(cells))
    cf = -np.ones((len(cells), 4), np.int64)
    cnt = np.zeros(len(cells), np.int64)
    for i in range(n):                                                # faces in ascending order, owner before neighbour
        c = cmap[owner[inter[i]]]
        cf[c, cnt[c]] = i; cnt[c] += 1
        if internal[i]:
            c = cmap[neighbour[inter[i]]]
            cf[c, cnt[c]] = i; cnt[c] += 1
    cp = ip[cf].reshape(len(cells), 8, 3)                             # (a face index of -1 picks the last face, as in the reference)
    tri = np.where(cnt == 3)[0]
    cp[tri, -2:, :] = 1e100
    with np.errstate(over="ignore", invalid="ignore"):
        d1 = np.linalg.norm(cp - cp[:, [0], :], axis=-1)[:, 1:].argmin(axis=1) + 1
        d2 = np.linalg.norm(cp - cp[:, [1], :], axis=-1)[:, 2:].argmin(axis=1) + 2
    d1 = d1 + (((d1 % 2 == 0) - 0.5) * 2).astype(np.int64)
    d2 = d2 + (((d2 % 2 == 0) - 0.5) * 2).astype(np.int64)
    ar = np.arange(len(cells))
    a = cp[ar, d1, :] - cp[:, 1, :]
    b = cp[ar, d2, :] - cp[:, 0, :]
    area = np.linalg.norm(np.cross(a, b), axis=1) / 2
    return cells.astype(np.int32), area
