"""adfvm_b200.blockmesh against the `boundary` files the reference ships next to its blockMeshDicts: the patch table
(nFaces / startFace of every patch, hence nInternalFaces) and the cell count must be exactly those of
cases/forwardStep/constant/polyMesh/boundary and cases/cylinder/constant/polyMesh/boundary (values restated here so that
the test runs without the reference tree), the mesh must be a valid hexahedral mesh for the metric build (positive volumes,
six faces per cell, matching cyclic planes)."""
import numpy as np

from adfvm_b200 import blockmesh
from adfvm_b200.metrics import build_mesh

# (patch, nFaces, startFace) as shipped
FORWARD_STEP = [("inlet", 80, 31936), ("outlet", 64, 32016), ("bottom", 48, 32080), ("top", 240, 32128), ("obstacle", 208, 32368),
                ("defaultFaces", 32256, 32576)]
CYLINDER = [("down", 300, 92000), ("right", 150, 92300), ("up", 300, 92450), ("left", 150, 92750), ("cylinder", 100, 92900),
            ("z1noc", 46250, 93000), ("z2noc", 46250, 139250)]


def _check(poly, table, ncells):
    assert [(k, v["nFaces"], v["startFace"]) for k, v in poly.boundary.items()] == table
    assert len(poly.neighbour) == table[0][2] and int(poly.owner.max()) + 1 == ncells
    assert len(poly.faces) == table[-1][1] + table[-1][2]
    assert np.all(poly.owner[:len(poly.neighbour)] < poly.neighbour)                  # upper-triangular order
    key = poly.owner[:len(poly.neighbour)].astype(np.int64) * ncells + poly.neighbour
    assert np.all(np.diff(key) > 0)
    m = build_mesh(poly)
    assert m.volumes.min() > 0 and m.areas.min() > 0
    return m


def test_forward_step_matches_shipped_boundary():
    m = _check(blockmesh.block_mesh(**blockmesh.forward_step_dict()), FORWARD_STEP, 16128)
    assert abs(m.volumes.sum() - (3 * 1 - 2.4 * 0.2) * 0.1) < 1e-12


def test_cylinder_matches_shipped_boundary():
    poly = blockmesh.block_mesh(**blockmesh.cylinder_dict())
    m = _check(poly, CYLINDER, 46250)
    b = poly.boundary
    f1 = m.faceCentres[b["z1noc"]["startFace"]:b["z1noc"]["startFace"] + 46250]
    f2 = m.faceCentres[b["z2noc"]["startFace"]:b["z2noc"]["startFace"] + 46250]
    assert np.abs(f1[:, :2] - f2[:, :2]).max() < 1e-15                                # face i of z1 faces face i of z2
    # the cylinder patch lies on the circle of radius 0.5 * 2.5e-4 (arc edges)
    fc = m.faceCentres[b["cylinder"]["startFace"]:b["cylinder"]["startFace"] + 100]
    r = np.linalg.norm(fc[:, :2], axis=1)
    assert np.abs(r / 1.25e-4 - 1).max() < 2e-4
    # half annulus + box: area of the domain in the x-y plane
    area = (13 * 5 - 0.5 * np.pi * 0.25) * 2.5e-4 ** 2
    assert abs(m.volumes.sum() / 2.5e-4 - area) / area < 1e-4
