"""OpenFOAM *binary* case files as the reference reads and writes them: the polyMesh
files `points, faces, owner, neighbour, boundary` (adFVM/mesh.py:206-279: regex header `FoamFile{... format binary ...}`,
then `N\\n(` raw little-endian bytes `)`; `faces` is a faceCompactList = int32 offsets (N+1) followed by int32 point
labels) and `volScalarField` / `volVectorField` files (adFVM/field.py:199-264, 411-476: `internalField nonuniform
List<scalar|vector> N (raw float64)` or `uniform ...`, one `boundaryField` dict per patch). Enough for a user of the
library to load a (decomposed: `processor<r>/`) case into `hexmesh.PolyMesh` + field arrays and to write results back
in the layout `apps/problem.py` / `apps/adjoint.py` leave on disk; also used by the golden-vector generator to hand
cases to the unmodified reference."""
import os
import re
from collections import OrderedDict

import numpy as np

from .hexmesh import PolyMesh

_HDR = """FoamFile
{{
    version     2.0;
    format      binary;
    class       {cls};
    location    "{loc}";
    object      {obj};
}}
"""


def time_name(t):
    """directory name of a time (adFVM/mesh.py:129-136): '%d' for integer times, else '%.11f'"""
    return "%d" % t if float(t).is_integer() else "%.11f" % t


# ------------------------------------------------------------------------------------------ writers
def _write_list(path, cls, loc, obj, arr):
    with open(path, "wb") as f:
        f.write(_HDR.format(cls=cls, loc=loc, obj=obj).encode())
        f.write(("\n%d\n(" % len(arr)).encode())
        f.write(arr.tobytes())
        f.write(b")\n")


def write_polymesh(case, poly):
    d = os.path.join(case, "constant", "polyMesh")
    os.makedirs(d, exist_ok=True)
    loc = "constant/polyMesh"
    _write_list(os.path.join(d, "points"), "vectorField", loc, "points", np.ascontiguousarray(poly.points, np.float64))
    _write_list(os.path.join(d, "owner"), "labelList", loc, "owner", np.ascontiguousarray(poly.owner, np.int32))
    _write_list(os.path.join(d, "neighbour"), "labelList", loc, "neighbour", np.ascontiguousarray(poly.neighbour, np.int32))
    nF = len(poly.faces)
    offsets = (4 * np.arange(nF + 1)).astype(np.int32)          # faceCompactList: offsets (nFaces+1), then point labels
    with open(os.path.join(d, "faces"), "wb") as f:
        f.write(_HDR.format(cls="faceCompactList", loc=loc, obj="faces").encode())
        f.write(("\n%d\n(" % (nF + 1)).encode()); f.write(offsets.tobytes()); f.write(b")\n")
        f.write(("\n%d\n(" % (4 * nF)).encode()); f.write(np.ascontiguousarray(poly.faces, np.int32).tobytes()); f.write(b")\n")
    with open(os.path.join(d, "boundary"), "w") as f:
        f.write(_HDR.format(cls="polyBoundaryMesh", loc=loc, obj="boundary"))
        f.write("\n%d\n(\n" % len(poly.boundary))
        for name, p in poly.boundary.items():
            f.write("    %s\n    {\n" % name)
            for k, v in p.items():
                if k.startswith("_") or k in ("cellStartFace",):
                    continue
                f.write("        %s %s;\n" % (k, v))
            f.write("    }\n")
        f.write(")\n")


def write_field(case, time_dir, name, internal, bfield):
    """internal: [nInternalCells, d]; bfield: {patch: {key: 'uniform ...' string | ndarray (nonuniform binary)}}"""
    d = os.path.join(case, time_dir)
    os.makedirs(d, exist_ok=True)
    internal = np.ascontiguousarray(internal, np.float64)
    vec = internal.shape[1] == 3
    with open(os.path.join(d, name), "wb") as f:
        f.write(_HDR.format(cls="volVectorField" if vec else "volScalarField", loc=time_dir, obj=name).encode())
        f.write(b"\ndimensions      [0 0 0 0 0 0 0];\n\n")
        f.write(("internalField   nonuniform List<%s> \n%d\n(" % ("vector" if vec else "scalar", len(internal))).encode())
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


# ------------------------------------------------------------------------------------------ readers
_HEADER = re.compile(rb"FoamFile\s*\{(.*?)\}", re.S)


def _after_header(data, path):
    m = _HEADER.search(data)
    if not m:
        raise ValueError("%s: no FoamFile header" % path)
    if not re.search(rb"format\s+binary\s*;", m.group(1)):
        raise ValueError("%s: only `format binary` files are supported (the reference's default, adFVM/config.py:198)" % path)
    return m.end()


def _binary_list(data, pos, dtype, width, path):
    """`N ( raw bytes )` starting at or after pos -> (array [N(,width)], position after the closing parenthesis)"""
    m = re.compile(rb"\s*(\d+)\s*\(").match(data, pos)
    if not m:
        raise ValueError("%s: list header not found" % path)
    n = int(m.group(1))
    nbytes = n * width * np.dtype(dtype).itemsize
    raw = data[m.end():m.end() + nbytes]
    if len(raw) != nbytes or data[m.end() + nbytes:m.end() + nbytes + 1] != b")":
        raise ValueError("%s: truncated binary list" % path)
    a = np.frombuffer(raw, dtype).copy()
    return (a.reshape(n, width) if width > 1 else a), m.end() + nbytes + 1


def _read_list(path, dtype, width=1):
    data = open(path, "rb").read()
    a, _ = _binary_list(data, _after_header(data, path), dtype, width, path)
    return a


def _parse_dict_entries(text):
    """`name { key value; ... }` blocks -> OrderedDict(name -> OrderedDict(key -> value string))"""
    out = OrderedDict()
    for m in re.finditer(r"([A-Za-z_][\w\-\.]*)\s*\{([^{}]*)\}", text):
        d = OrderedDict()
        for kv in re.finditer(r"([A-Za-z_]\w*)\s+([^;]+);", m.group(2)):
            d[kv.group(1)] = kv.group(2).strip()
        out[m.group(1)] = d
    return out


def read_boundary(case):
    path = os.path.join(case, "constant", "polyMesh", "boundary")
    data = open(path, "rb").read()
    text = data[_HEADER.search(data).end():].decode()
    text = text[text.index("(") + 1:text.rindex(")")]
    boundary = OrderedDict()
    for name, d in _parse_dict_entries(text).items():
        for k in ("nFaces", "startFace", "myProcNo", "neighbProcNo"):
            if k in d:
                d[k] = int(d[k])
        boundary[name] = d
    return boundary


def read_polymesh(case):
    """constant/polyMesh of `case` (or of case/processor<r>) -> hexmesh.PolyMesh (quad faces only, like the reference)"""
    d = os.path.join(case, "constant", "polyMesh")
    points = _read_list(os.path.join(d, "points"), np.float64, 3)
    owner = _read_list(os.path.join(d, "owner"), np.int32)
    neighbour = _read_list(os.path.join(d, "neighbour"), np.int32)
    path = os.path.join(d, "faces")
    data = open(path, "rb").read()
    offsets, pos = _binary_list(data, _after_header(data, path), np.int32, 1, path)
    labels, _ = _binary_list(data, pos, np.int32, 1, path)
    if not np.all(np.diff(offsets) == 4):
        raise ValueError("%s: hexahedral meshes only (every face a quad, adFVM/mesh.py:216-225)" % path)
    faces = labels.reshape(-1, 4)
    if len(faces) != len(owner):
        raise ValueError("faces / owner size mismatch")
    return PolyMesh(points, faces, owner, neighbour, read_boundary(case))


def read_field(case, time_dir, name, nInternalCells, boundary):
    """-> (internal [nInternalCells, d] float64, {patch: {key: str | ndarray}}); `uniform` internal fields are expanded"""
    path = os.path.join(case, time_dir, name)
    data = open(path, "rb").read()
    pos = _after_header(data, path)
    hdr = _HEADER.search(data).group(1)
    vec = re.search(rb"class\s+volVectorField\s*;", hdr) is not None
    width = 3 if vec else 1
    m = re.compile(rb"internalField\s+(uniform|nonuniform)\s*").search(data, pos)
    if not m:
        raise ValueError("%s: no internalField" % path)
    if m.group(1) == b"uniform":
        e = data.index(b";", m.end())
        toks = data[m.end():e].replace(b"(", b" ").replace(b")", b" ").split()
        try:
            vals = [float(x) for x in toks]
        except ValueError:
            raise ValueError("%s: cannot parse uniform internalField %r" % (path, data[m.end():e]))
        if len(vals) != width:
            raise ValueError("%s: uniform internalField has %d components, expected %d" % (path, len(vals), width))
        internal = np.tile(np.array(vals, np.float64).reshape(1, width), (nInternalCells, 1))
        pos = e + 1
    else:
        lm = re.compile(rb"List<(scalar|vector)>\s*").match(data, m.end())
        internal, pos = _binary_list(data, lm.end(), np.float64, width, path)
        internal = internal.reshape(-1, width)
        if len(internal) != nInternalCells:
            raise ValueError("%s: %d values for %d cells" % (path, len(internal), nInternalCells))
    # boundaryField { patch { key value; ... } ... }: parsed sequentially, block by block; binary lists are skipped by their
    # declared length, so that payload bytes can never be mistaken for a patch name or a key
    bm = re.compile(rb"\s*boundaryField\s*\{").search(data, pos)
    if not bm:
        raise ValueError("%s: no boundaryField" % path)
    p = bm.end()
    found = OrderedDict()
    name_re = re.compile(rb"\s*(\}|[^\s{};]+)\s*")
    while True:
        nm_ = name_re.match(data, p)
        if not nm_:
            raise ValueError("%s: malformed boundaryField" % path)
        if nm_.group(1) == b"}":
            break
        pname, p = nm_.group(1).decode(), nm_.end()
        if data[p:p + 1] != b"{":
            raise ValueError("%s: expected '{' after patch name %s" % (path, pname))
        p += 1
        d = OrderedDict()
        while True:
            km = re.compile(rb"\s*(\}|[A-Za-z_]\w*)").match(data, p)
            if not km:
                raise ValueError("%s: malformed entry in patch %s" % (path, pname))
            if km.group(1) == b"}":
                p = km.end()
                break
            key, p = km.group(1).decode(), km.end()
            nm = re.compile(rb"\s+nonuniform\s+List<(scalar|vector)>\s*").match(data, p)
            if nm:
                arr, p = _binary_list(data, nm.end(), np.float64, 3 if nm.group(1) == b"vector" else 1, path)
                d[key] = arr.reshape(len(arr), -1)
                p = data.index(b";", p) + 1
            else:
                e = data.index(b";", p)
                d[key] = data[p:e].decode().strip()
                p = e + 1
        found[pname] = d
    bfield = OrderedDict()
    for patch in boundary:
        if patch not in found:
            raise ValueError("%s: no boundaryField entry for patch %s" % (path, patch))
        bfield[patch] = found[patch]
    return internal, bfield
