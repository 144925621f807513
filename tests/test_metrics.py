"""adfvm_b200.metrics (numpy restatement of cmesh.cpp) against the mesh arrays the reference's own
cmesh.build produced for the golden cases (positional inputs 4..18 of the recorded `primal` calls)."""
import numpy as np
import pytest

from golden_util import Golden, available
from adfvm_b200.hexmesh import PolyMesh
from adfvm_b200.metrics import build_mesh, GRAD_FIELDS, INT_FIELDS

CASES = [c for c in available() if not c.endswith("_fp32")]


def rebuild(g):
    spec = g.spec
    points, faces = g.mesh_array("orig", "points"), g.mesh_array("orig", "faces")[:, 1:]
    _, _, inp, _, _ = next(g.calls("orig", "primal"))
    owner, nIF = inp[14], inp[22]
    from collections import OrderedDict
    boundary = OrderedDict()
    for p in sorted(spec["patches"], key=lambda p: p["startFace"]):
        t = "patch" if p["type"] == "characteristic" else p["type"]
        d = dict(type=t, nFaces=p["nFaces"], startFace=p["startFace"])
        if "neighbourPatch" in p:
            d["neighbourPatch"] = p["neighbourPatch"]
        boundary[p["name"]] = d
    poly = PolyMesh(points, faces, owner, inp[15][:nIF], boundary)
    return build_mesh(poly), inp


@pytest.mark.parametrize("name", CASES)
def test_metrics_match_reference(name):
    g = Golden(name)
    m, inp = rebuild(g)
    for k, a in enumerate(GRAD_FIELDS):
        ref = inp[4 + k]
        mine = getattr(m, a)
        assert mine.shape == ref.shape, a
        np.testing.assert_allclose(mine, ref, rtol=1e-12, atol=1e-14 * np.abs(ref).max(), err_msg=a)
    for k, a in enumerate(INT_FIELDS):
        assert np.array_equal(getattr(m, a), inp[14 + k]), a
    assert m.getScalar() == [int(x) for x in inp[19:19 + 8 + 3 * len(m.sortedPatches)]]
    np.testing.assert_allclose(m.cellCentres, g.mesh_array("orig", "cellCentres"), rtol=1e-12, atol=1e-14)


def test_uniform_box_matches_general_builder():
    """the closed-form box generator used for benchmark-size meshes == the general metric builder"""
    import numpy as np
    from adfvm_b200 import hexmesh, metrics
    walls = [("inlet", "patch", ["x-"], {}), ("outlet", "patch", ["x+"], {}), ("floor", "symmetryPlane", ["y-"], {}),
             ("lid", "patch", ["y+"], {}), ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}),
             ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})]
    for n, lo, hi, patches in (((5, 4, 3), (0, 0, 0), (1, 2, 0.5), None), ((4, 3, 5), (0.5, 0, 0), (2, 1, 0.5), walls)):
        a = metrics.build_mesh(hexmesh.box_mesh(n, lo, hi, patches=patches))
        b = metrics.uniform_box(n, lo, hi, patches)
        for k in metrics.GRAD_FIELDS + ["cellCentres"]:
            x, y = getattr(a, k), getattr(b, k)
            assert x.shape == y.shape and np.abs(x - y).max() <= 1e-12 * max(1, np.abs(x).max()), k
        for k in metrics.INT_FIELDS:
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
        assert a.getScalar() == b.getScalar() and a.sortedPatches == b.sortedPatches
