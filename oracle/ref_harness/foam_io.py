"""TEST INFRASTRUCTURE ONLY — writes a PolyMesh + initial fields as an OpenFOAM *binary* case that
the unmodified reference can read (formats restated from the reference's regex-based readers:
adFVM/mesh.py:206-279 for polyMesh files, adFVM/field.py:199-264 for fields)."""
import os

import numpy as np

_HDR = """FoamFile
{{
    version     2.0;
    format      binary;
    class       {cls};
    location    "{loc}";
    object      {obj};
}}
"""


def _write_list(path, cls, loc, obj, arr):
    with open(path, "wb") as f:
        f.write(_HDR.format(cls=cls, loc=loc, obj=obj).encode())
        f.write(("\n%d\n(" % len(arr)).encode())
        f.write(arr.tobytes())
        f.write(b")\n")


def write_polymesh(case, poly):
    d = os.path.join(case, "constant", "polyMesh")
    os.makedirs(d, exist_ok=True)
    _write_list(os.path.join(d, "points"), "vectorField", "constant/polyMesh", "points",
                np.ascontiguousarray(poly.points, np.float64))
    _write_list(os.path.join(d, "owner"), "labelList", "constant/polyMesh", "owner",
                np.ascontiguousarray(poly.owner, np.int32))
    _write_list(os.path.join(d, "neighbour"), "labelList", "constant/polyMesh", "neighbour",
                np.ascontiguousarray(poly.neighbour, np.int32))
    # faceCompactList: offsets (nFaces+1) then point labels
    nF = len(poly.faces)
    offsets = (4 * np.arange(nF + 1)).astype(np.int32)
    with open(os.path.join(d, "faces"), "wb") as f:
        f.write(_HDR.format(cls="faceCompactList", loc="constant/polyMesh", obj="faces").encode())
        f.write(("\n%d\n(" % (nF + 1)).encode())
        f.write(offsets.tobytes())
        f.write(b")\n")
        f.write(("\n%d\n(" % (4 * nF)).encode())
        f.write(np.ascontiguousarray(poly.faces, np.int32).tobytes())
        f.write(b")\n")
    with open(os.path.join(d, "boundary"), "w") as f:
        f.write(_HDR.format(cls="polyBoundaryMesh", loc="constant/polyMesh", obj="boundary"))
        f.write("\n%d\n(\n" % len(poly.boundary))
        for name, p in poly.boundary.items():
            f.write("    %s\n    {\n" % name)
            for k, v in p.items():
                if k.startswith("_") or k in ("cellStartFace",):
                    continue
                f.write("        %s %s;\n" % (k, v))
            f.write("    }\n")
        f.write(")\n")


def write_field(case, time_name, name, internal, bfield):
    """internal: [nInternalCells, d] array; bfield: dict patch -> dict(type=..., key=value...) where
    value is a string (uniform ...) or an ndarray (written as nonuniform binary)."""
    d = os.path.join(case, time_name)
    os.makedirs(d, exist_ok=True)
    internal = np.ascontiguousarray(internal, np.float64)
    vec = internal.shape[1] == 3
    cls = "volVectorField" if vec else "volScalarField"
    typ = "vector" if vec else "scalar"
    with open(os.path.join(d, name), "wb") as f:
        f.write(_HDR.format(cls=cls, loc=time_name, obj=name).encode())
        f.write(b"\ndimensions      [0 0 0 0 0 0 0];\n\n")
        f.write(("internalField   nonuniform List<%s> \n%d\n(" % (typ, len(internal))).encode())
        f.write(internal.tobytes())
        f.write(b")\n;\n\nboundaryField\n{\n")
        for patch, dd in bfield.items():
            f.write(("    %s\n    {\n" % patch).encode())
            for k, v in dd.items():
                if isinstance(v, np.ndarray):
                    v = np.ascontiguousarray(v, np.float64)
                    t = "vector" if (v.ndim == 2 and v.shape[1] == 3) else "scalar"
                    f.write(("        %s nonuniform List<%s> \n%d\n(" % (k, t, len(v))).encode())
                    f.write(v.tobytes())
                    f.write(b")\n;\n")
                else:
                    f.write(("        %s %s;\n" % (k, v)).encode())
            f.write(b"    }\n")
        f.write(b"}\n")
