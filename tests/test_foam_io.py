"""OpenFOAM binary case files (adfvm_b200/foam_io.py): write -> read round trips of polyMesh and of
fields with uniform / nonuniform boundary values. That the unmodified reference reads what the writer produces is what
every golden fixture rests on (oracle/ref_harness/gen_golden.py writes its cases with it)."""
import numpy as np

from adfvm_b200 import foam_io, hexmesh
from adfvm_b200.metrics import build_mesh


def test_polymesh_and_field_round_trip(tmp_path):
    lo, hi = (0., 0., 0.), (2., 1., .5)
    poly = hexmesh.box_mesh((5, 4, 3), lo, hi, grading=(1.0, 0.4, 1.0), warp=hexmesh.sine_warp(0.02, lo, hi), patches=[
        ("inlet", "patch", ["x-"], {}), ("outlet", "patch", ["x+"], {}), ("floor", "symmetryPlane", ["y-"], {}),
        ("lid", "patch", ["y+"], {}), ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}), ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    case = str(tmp_path / "processor0")                       # decomposed-case layout: processor<r>/constant/polyMesh
    foam_io.write_polymesh(case, poly)
    back = foam_io.read_polymesh(case)
    assert np.array_equal(back.points, poly.points) and np.array_equal(back.faces, poly.faces)
    assert np.array_equal(back.owner, poly.owner) and np.array_equal(back.neighbour, poly.neighbour)
    assert list(back.boundary) == list(poly.boundary)
    for k in poly.boundary:
        for key in ("type", "nFaces", "startFace"):
            assert back.boundary[k][key] == poly.boundary[k][key]
    assert back.boundary["z1"]["neighbourPatch"] == "z2"
    a, b = build_mesh(poly), build_mesh(back)
    assert np.array_equal(a.volumes, b.volumes) and np.array_equal(a.quadraticWeights, b.quadraticWeights)
    n = a.nInternalCells
    rng = np.random.RandomState(0)
    U = rng.randn(n, 3); T = 300 + rng.rand(n, 1)
    nl = poly.boundary["lid"]["nFaces"]
    bU = {"inlet": {"type": "calculated"}, "outlet": {"type": "zeroGradient"}, "floor": {"type": "symmetryPlane"},
          "lid": {"type": "fixedValue", "value": rng.randn(nl, 3)}, "z1": {"type": "cyclic"}, "z2": {"type": "cyclic"}}
    bT = {"inlet": {"type": "CBC_TOTAL_PT", "Tt": "uniform 305", "pt": rng.rand(poly.boundary["inlet"]["nFaces"], 1)},
          "outlet": {"type": "zeroGradient"}, "floor": {"type": "symmetryPlane"}, "lid": {"type": "fixedValue", "value": "uniform 310"},
          "z1": {"type": "cyclic"}, "z2": {"type": "cyclic"}}
    t = foam_io.time_name(2e-6)
    assert t == "0.00000200000" and foam_io.time_name(3.0) == "3"
    foam_io.write_field(case, t, "U", U, bU)
    foam_io.write_field(case, t, "T", T, bT)
    Ui, bUr = foam_io.read_field(case, t, "U", n, poly.boundary)
    Ti, bTr = foam_io.read_field(case, t, "T", n, poly.boundary)
    assert np.array_equal(Ui, U) and np.array_equal(Ti, T)
    assert np.array_equal(bUr["lid"]["value"], bU["lid"]["value"]) and bUr["lid"]["type"] == "fixedValue"
    assert bTr["inlet"]["Tt"] == "uniform 305" and np.array_equal(bTr["inlet"]["pt"], bT["inlet"]["pt"])
    assert bTr["lid"]["value"] == "uniform 310" and bUr["z1"]["type"] == "cyclic"


def test_uniform_internal_field(tmp_path):
    poly = hexmesh.box_mesh((2, 2, 2))
    case = str(tmp_path)
    foam_io.write_polymesh(case, poly)
    path = tmp_path / "0"
    path.mkdir()
    (path / "p").write_bytes(b'FoamFile\n{\n    version 2.0;\n    format binary;\n    class volScalarField;\n    object p;\n}\n'
                             b'dimensions [1 -1 -2 0 0 0 0];\ninternalField   uniform 101325;\nboundaryField\n{\n' +
                             b"".join(("    %s\n    {\n        type cyclic;\n    }\n" % k).encode() for k in poly.boundary) + b"}\n")
    internal, bf = foam_io.read_field(case, "0", "p", 8, poly.boundary)
    assert internal.shape == (8, 1) and np.all(internal == 101325.) and bf["x1"]["type"] == "cyclic"
