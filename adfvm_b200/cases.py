"""Synthetic cases in the reference's own input format: mesh arrays, initial fields, the static `spec`
and the positional argument list `solver.map(*inputs)` would receive (adFVM/solver.py:312-317).

Used by the parity tests, `__graft_entry__.smoke()` and `bench.py` (hex box [0,1]^3, N^3 cells,
cyclic in x,y,z, smooth sinusoidal U/T/p, Gaussian source perturbation, objective sum T*V), plus a walled
variant exercising the other boundary conditions.
"""
from __future__ import annotations

import numpy as np

from . import hexmesh
from .metrics import build_mesh, uniform_box, GRAD_FIELDS, INT_FIELDS


class Case:
    """mesh: metrics.MeshData; spec: dict; fields: (rho, rhoU, rhoE) conservative initial state;
    source: 3 arrays; bcvals: {(field, patch, key): array}"""

    def __init__(self, mesh, spec, state, source, bcvals, dt, dtype=np.float64):
        self.mesh, self.spec, self.dt = mesh, spec, dt
        self.dtype = np.dtype(dtype)
        c = lambda a: np.ascontiguousarray(a, self.dtype)
        self.state = [c(s) for s in state]
        self.source = [c(s) for s in source]
        self.bcvals = {k: c(v) for k, v in bcvals.items()}
        self._static = None

    def static_inputs(self):
        if self._static is None:
            m = self.mesh
            t = [np.ascontiguousarray(getattr(m, a), self.dtype) for a in GRAD_FIELDS] + \
                [np.ascontiguousarray(getattr(m, a), np.int32) for a in INT_FIELDS]
            bc = []
            for field in ("U", "T", "p"):
                for pid in self.spec["sortedPatches"]:
                    for key in self.spec["BCs"][field][pid]["keys"]:
                        bc.append(self.bcvals[(field, pid, key)])
            self._static = (t + m.getScalar(), self.source, bc)
        return self._static

    def inputs(self, state=None, dt=None):
        """positional list for `primal` (extraArgs of the objective - `self.extra`, if any - after the BC arrays)"""
        mesh_args, source, bc = self.static_inputs()
        st = self.state if state is None else state
        return list(st) + [np.array([[self.dt if dt is None else dt]], self.dtype)] + mesh_args + list(source) + bc + \
            list(getattr(self, "extra", []))

    def adjoint_inputs(self, state, adj, obja=1.0, dtca=0.0, scaling=0.0, dt=None):
        """positional list for `primal_grad` (apps/adjoint.py:272-280)"""
        a = lambda v: np.array([[v]], self.dtype)
        return self.inputs(state, dt) + list(adj) + [a(dtca), a(obja)] + [a(scaling)]


def conservative(U, T, p, gamma=1.4, Cp=1004.5):
    """adFVM/density.py:173-181"""
    Cv = Cp / gamma
    e = Cv * T
    rho = p / (e * (gamma - 1))
    rhoE = rho * (e + 0.5 * (U * U).sum(axis=1, keepdims=True))
    return rho, U * rho, rhoE


def _spec(mesh, bcs, objective, mu=None, briemann="eulerRoe", Cp=1004.5, gamma=1.4, Pr=0.7):
    patches = []
    for pid in mesh.sortedPatches + mesh.remotePatches:
        p = mesh.boundary[pid]
        d = {"name": pid, "type": p["type"], "startFace": p["startFace"], "nFaces": p["nFaces"],
             "cellStartFace": p["cellStartFace"]}
        for k in ("neighbourPatch", "myProcNo", "neighbProcNo", "referPatch", "tag"):
            if k in p:
                d[k] = p[k]
        patches.append(d)
    return {"Cp": Cp, "gamma": gamma, "Pr": Pr, "mu": mu or {"law": "sutherland"}, "riemannSolver": "eulerRoe",
            "boundaryRiemannSolver": briemann, "timeIntegrator": "SSPRK", "patches": patches, "BCs": bcs,
            "sortedPatches": list(mesh.sortedPatches), "objective": objective}


def smooth_state(cc, lo=(0., 0., 0.), hi=(1., 1., 1.)):
    """s = sin2πx·cos2πy·sin2πz; U=(100+10s, 50−5s, 20+2s), T=300+10s, p=101325+1000s"""
    s3 = (cc - np.asarray(lo)) / (np.asarray(hi) - np.asarray(lo))
    s = np.sin(2 * np.pi * s3[:, 0]) * np.cos(2 * np.pi * s3[:, 1]) * np.sin(2 * np.pi * s3[:, 2])
    U = np.stack([100 + 10 * s, 50 - 5 * s, 20 + 2 * s], axis=1)
    return U, (300 + 10 * s).reshape(-1, 1), (101325 + 1000 * s).reshape(-1, 1)


def gaussian_source(cc, mid=(0.5, 0.5, 0.5), amp=1e2, width=50.):
    """G = 1e2·exp(−50|x−½|²) → (G, (100G,0,0), 2e5·G)"""
    G = amp * np.exp(-width * ((cc - np.asarray(mid)) ** 2).sum(axis=1, keepdims=True))
    rhoU = np.zeros((len(cc), 3)); rhoU[:, 0] = 100 * G[:, 0]
    return G, rhoU, 2e5 * G


def periodic_box(n, dtype=np.float64, warp=0.0, dt=None, mesh=None):
    """The §8(d) benchmark workload: n^3 (or (nx,ny,nz)) periodic unit box."""
    if isinstance(n, int):
        n = (n, n, n)
    if mesh is None:
        mesh = build_mesh(hexmesh.box_mesh(n, warp=hexmesh.sine_warp(warp))) if warp else uniform_box(n)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    U, T, p = smooth_state(cc)
    bcs = {f: {pid: {"type": "cyclic", "keys": []} for pid in mesh.sortedPatches} for f in ("U", "T", "p")}
    spec = _spec(mesh, bcs, {"kind": "cell_TV"})
    if dt is None:
        dt = 1e-6 * 48. / max(n)        # dt = 1e-6 at 48^3, scaled ∝ 1/N
    return Case(mesh, spec, conservative(U, T, p), gaussian_source(cc), {}, dt, dtype)


def shock_tube(n=500, width=0.06, dtype=np.float64, dt=1e-5):
    """BASELINE.json config 1: the geometry of the reference's cases/shockTube (n x 1 x 1 cells on [-5,5] x [-1,1]^2, patches
    `sides` = {x=+5, x=-5} zeroGradient, `empty` = the 4n lateral faces), tanh-smoothed Sod initial condition
    sigma = (1 - tanh(x/width))/2 (the shipped sharp one NaNs in the reference), inviscid, objective
    sum_{sides} p_ghost * area, source perturbation G = 1e3 exp(-((x+4.5)/0.2)^2) (BASELINE.md §5 anchor B)."""
    lo, hi = (-5., -1., -1.), (5., 1., 1.)
    poly = hexmesh.box_mesh((n, 1, 1), lo, hi, patches=[("sides", "patch", ["x+", "x-"], {}),
                                                        ("empty", "empty", ["y-", "z+", "y+", "z-"], {})])
    mesh = build_mesh(poly)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    sig = 0.5 * (1 - np.tanh(cc[:, 0] / width))
    p = (1e4 + 9e4 * sig).reshape(-1, 1)
    rho = (0.125 + 0.875 * sig).reshape(-1, 1)
    R = 1004.5 - 1004.5 / 1.4
    T = p / (rho * R)
    U = np.zeros((C, 3))
    k0 = {"keys": []}
    bcs = {f: {"sides": dict(type="zeroGradient", **k0), "empty": dict(type="zeroGradient", **k0)} for f in ("U", "T", "p")}   # class names, as the reference reports them (empty = zeroGradient, BCs.py)
    spec = _spec(mesh, bcs, {"kind": "patch_pA", "patch": "sides"}, mu={"law": "constant", "value": 0.})
    return Case(mesh, spec, conservative(U, T, p), gaussian_source(cc, (-4.5, 0., 0.), 1e3, 25.), {}, dt, dtype)


def walled_box(n=(8, 6, 4), dtype=np.float64, warp=0.02, dt=2e-6, nonuniform=True):
    """Channel exercising every supported BC: CBC_TOTAL_PT inlet, fixedValue-p outlet, symmetryPlane floor,
    no-slip isothermal lid (fixedValue U,T), one cyclic pair; constant viscosity; drag objective on the lid.
    With nonuniform=True the BC value arrays vary face by face (the reference's readers cannot express
    that — SURVEY H6 — so it is checked against the oracle only)."""
    lo, hi = (0., 0., 0.), (2., 1., 0.5)
    poly = hexmesh.box_mesh(n, lo, hi, grading=(1.0, 0.4, 1.0), warp=hexmesh.sine_warp(warp, lo, hi) if warp else None,
                            patches=[("inlet", "patch", ["x-"], {}), ("outlet", "patch", ["x+"], {}),
                                     ("floor", "symmetryPlane", ["y-"], {}), ("lid", "patch", ["y+"], {}),
                                     ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}),
                                     ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    mesh = build_mesh(poly)
    mesh.boundary["inlet"]["type"] = "characteristic"       # what CBC_TOTAL_PT.__init__ does (BCs.py:167)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    s3 = (cc - np.asarray(lo)) / (np.asarray(hi) - np.asarray(lo))
    s = np.sin(2 * np.pi * s3[:, 0]) * np.cos(np.pi * s3[:, 1]) * np.cos(2 * np.pi * s3[:, 2])
    U = np.stack([60 * (1 - s3[:, 1] ** 2) + 3 * s, 2 * s, 1 * s], axis=1)
    T = (300 + 5 * s).reshape(-1, 1)
    p = (101325 + 500 * s).reshape(-1, 1)
    cyc = {"z1": {"type": "cyclic", "keys": []}, "z2": {"type": "cyclic", "keys": []}}
    k0 = {"keys": []}
    bcs = {"U": dict(cyc, inlet=dict(type="calculated", **k0), outlet=dict(type="zeroGradient", **k0),
                     floor=dict(type="symmetryPlane", **k0), lid={"type": "fixedValue", "keys": ["value"]}),
           "T": dict(cyc, inlet=dict(type="calculated", **k0), outlet=dict(type="zeroGradient", **k0),
                     floor=dict(type="symmetryPlane", **k0), lid={"type": "fixedValue", "keys": ["value"]}),
           "p": dict(cyc, inlet={"type": "CBC_TOTAL_PT", "keys": ["Tt", "pt"]},
                     outlet={"type": "fixedValue", "keys": ["value"]},
                     floor=dict(type="symmetryPlane", **k0), lid=dict(type="zeroGradient", **k0))}
    nin, nout, nlid = (mesh.boundary[k]["nFaces"] for k in ("inlet", "outlet", "lid"))
    v = (lambda n_: np.sin(1.0 + np.arange(n_))) if nonuniform else (lambda n_: np.zeros(n_))
    Ulid = np.zeros((nlid, 3)); Ulid[:, 0] = 1.5 + 0.2 * v(nlid)
    bcvals = {("U", "lid", "value"): Ulid, ("T", "lid", "value"): (310 + 2 * v(nlid)).reshape(-1, 1),
              ("p", "outlet", "value"): (101000 + 50 * v(nout)).reshape(-1, 1),
              ("p", "inlet", "Tt"): (305 + v(nin)).reshape(-1, 1), ("p", "inlet", "pt"): (103000 + 100 * v(nin)).reshape(-1, 1)}
    spec = _spec(mesh, bcs, {"kind": "drag", "patch": "lid", "direction": 0}, mu={"law": "constant", "value": 2.5e-5})
    return Case(mesh, spec, conservative(U, T, p), gaussian_source(cc, (1.0, 0.5, 0.25), 1e2, 20.), bcvals, dt, dtype)


def cylinder2d(nr=185, nt=250, dtype=np.float64, dt=2e-9):
    """2-D laminar cylinder at the size of the reference's cases/cylinder mesh (C = 46 250 cells, BASELINE.json config 2;
    the shipped blockMeshDict cannot be meshed here and its outlet BC has no class in BCs.py): half annulus around a
    cylinder of radius 0.5 mm, geometric radial spacing, one cell in the span with cyclic z1/z2, no-slip wall,
    CBC_UPT far field with the Lax-Friedrichs boundary solver, symmetry planes on the axis, mu = 2.5e-5, drag
    objective (templates/cylinder.py:9-19). Same set-up as the golden fixture `cyl2d` (12 x 16 cells)."""
    r0, r1 = 0.5e-3, 8e-3

    def warp(p):
        r = r0 * (r1 / r0) ** p[:, 0]
        th = np.pi * p[:, 1]
        return np.stack([r * np.cos(th), r * np.sin(th), p[:, 2]], axis=1)
    poly = hexmesh.box_mesh((nr, nt, 1), (0., 0., 0.), (1., 1., 2e-4), warp=warp, patches=[
        ("cylinder", "patch", ["x-"], {}), ("far", "patch", ["x+"], {}), ("axis", "symmetryPlane", ["y-", "y+"], {}),
        ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}), ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    mesh = build_mesh(poly)
    mesh.boundary["far"]["type"] = "characteristic"          # what CBC_UPT.__init__ does (BCs.py:142)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    r = np.linalg.norm(cc[:, :2], axis=1)
    th = np.arctan2(cc[:, 1], cc[:, 0])
    U0 = 33.
    f = 1 - (r0 / r) ** 2
    U = np.stack([U0 * (1 - (r0 / r) ** 2 * np.cos(2 * th)) * f, -U0 * (r0 / r) ** 2 * np.sin(2 * th) * f, 0 * r], axis=1)
    T = (300 + 2 * np.exp(-((r - r0) / 1e-3) ** 2)).reshape(-1, 1)
    p = (101325 + 0.5 * 1.17 * (U0 ** 2 - (U ** 2).sum(axis=1))).reshape(-1, 1)
    k0 = {"keys": []}
    cyc = {"z1": dict(type="cyclic", **k0), "z2": dict(type="cyclic", **k0)}
    bcs = {"U": dict(cyc, cylinder={"type": "fixedValue", "keys": ["value"]}, far=dict(type="calculated", **k0), axis=dict(type="symmetryPlane", **k0)),
           "T": dict(cyc, cylinder=dict(type="zeroGradient", **k0), far=dict(type="calculated", **k0), axis=dict(type="symmetryPlane", **k0)),
           "p": dict(cyc, cylinder=dict(type="zeroGradient", **k0), far={"type": "CBC_UPT", "keys": ["U0", "T0", "p0"]}, axis=dict(type="symmetryPlane", **k0))}
    ncyl, nfar = mesh.boundary["cylinder"]["nFaces"], mesh.boundary["far"]["nFaces"]
    Ufar = np.zeros((nfar, 3)); Ufar[:, 0] = U0
    bcvals = {("U", "cylinder", "value"): np.zeros((ncyl, 3)), ("p", "far", "U0"): Ufar,
              ("p", "far", "T0"): np.full((nfar, 1), 300.), ("p", "far", "p0"): np.full((nfar, 1), 101325.)}
    spec = _spec(mesh, bcs, {"kind": "drag", "patch": "cylinder", "direction": 0}, mu={"law": "constant", "value": 2.5e-5},
                 briemann="eulerLaxFriedrichs")
    return Case(mesh, spec, conservative(U, T, p), gaussian_source(cc, (-0.8e-3, 0.1e-3, 1e-4), 1e-1, 2.5e6), bcvals, dt, dtype)


def forward_step(nx=240, ny=80, dtype=np.float64, dt=1e-4):
    """Mach-3 forward-facing step at the size of the reference's cases/forwardStep mesh (C = 16 128 cells, BASELINE.json
    config 3): box [0,3]x[0,1] minus the step x > 0.6, y < 0.2, one cell in the span with `empty` faces, Cp = 2.5,
    inviscid, fixedValue inlet, inletOutlet (= zeroGradient, BCs.py:137) outlet, symmetryPlane top/bottom, slip
    (= symmetryPlane, BCs.py:135) obstacle, pressure-force objective on the obstacle (templates/forwardStep.py); BC types
    are given by the names of the classes they resolve to, as the reference's own objects report them. Same set-up as the golden fixture `step2d` (30 x 10 cells)."""
    lo, hi = (0., 0., -0.05), (3., 1., 0.05)
    K, J, I = np.meshgrid(np.arange(1), np.arange(ny), np.arange(nx), indexing="ij")
    keep = ~((I >= nx // 5) & (J < ny // 5))
    poly = hexmesh.masked_box_mesh((nx, ny, 1), lo, hi, keep, [
        ("inlet", "patch", ["x-"], {}), ("outlet", "patch", ["x+"], {}), ("bottom", "symmetryPlane", ["y-"], {}),
        ("top", "symmetryPlane", ["y+"], {}), ("defaultFaces", "empty", ["z-", "z+"], {})], hole=("obstacle", "patch", {}))
    mesh = build_mesh(poly)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    s = np.sin(2 * np.pi * cc[:, 0] / 3.) * np.cos(np.pi * cc[:, 1])
    U = np.stack([3. + 0.05 * s, 0.02 * s, 0 * s], axis=1)
    T = (1. + 0.01 * s).reshape(-1, 1)
    p = (1. + 0.02 * s).reshape(-1, 1)
    k0 = {"keys": []}
    fv = {"type": "fixedValue", "keys": ["value"]}
    bcs = {"U": {"inlet": fv, "outlet": dict(type="zeroGradient", **k0), "bottom": dict(type="symmetryPlane", **k0),
                 "top": dict(type="symmetryPlane", **k0), "obstacle": dict(type="symmetryPlane", **k0), "defaultFaces": dict(type="zeroGradient", **k0)},
           "T": {"inlet": fv, "outlet": dict(type="zeroGradient", **k0), "bottom": dict(type="symmetryPlane", **k0),
                 "top": dict(type="symmetryPlane", **k0), "obstacle": dict(type="zeroGradient", **k0), "defaultFaces": dict(type="zeroGradient", **k0)},
           "p": {"inlet": fv, "outlet": dict(type="zeroGradient", **k0), "bottom": dict(type="symmetryPlane", **k0),
                 "top": dict(type="symmetryPlane", **k0), "obstacle": dict(type="zeroGradient", **k0), "defaultFaces": dict(type="zeroGradient", **k0)}}
    nin = mesh.boundary["inlet"]["nFaces"]
    Uin = np.zeros((nin, 3)); Uin[:, 0] = 3.
    bcvals = {("U", "inlet", "value"): Uin, ("T", "inlet", "value"): np.ones((nin, 1)), ("p", "inlet", "value"): np.ones((nin, 1))}
    spec = _spec(mesh, bcs, {"kind": "patch_pA", "patch": "obstacle"}, mu={"law": "constant", "value": 0.}, Cp=2.5)
    return Case(mesh, spec, conservative(U, T, p, Cp=2.5), gaussian_source(cc, (0.3, 0.3, 0.), 1e-2, 20.), bcvals, dt, dtype)


# ------------------------------------------------------------------------------------------ the shipped cases on their own meshes
def forward_step_shipped(dtype=np.float64, dt=1e-4):
    """BASELINE.json config 3 as the reference ships it: the mesh of cases/forwardStep/constant/polyMesh/blockMeshDict (3 blocks,
    16 128 cells, generated by adfvm_b200.blockmesh; patch table = the shipped `boundary` file), the shipped initial fields
    (cases/forwardStep/0: uniform U = (3,0,0), T = 1, p = 1) and the set-up of templates/forwardStep.py: Cp = 2.5, inviscid,
    fixedValue inlet, inletOutlet outlet (= zeroGradient, BCs.py:137), symmetryPlane top / bottom, slip obstacle, `empty` span,
    objective sum p_ghost * area over the obstacle. Source perturbation: a Gaussian just upstream of the step face (the template's
    momentum source in the inlet cells, :32-40, cannot reach the obstacle within the 20 steps of the recorded anchor: its
    sensitivity is exactly zero there)."""
    from . import blockmesh
    mesh = build_mesh(blockmesh.block_mesh(**blockmesh.forward_step_dict()))
    C = mesh.nInternalCells
    U = np.zeros((C, 3)); U[:, 0] = 3.
    T = np.ones((C, 1)); p = np.ones((C, 1))
    k0 = {"keys": []}
    fv = {"type": "fixedValue", "keys": ["value"]}
    zg, sym = dict(type="zeroGradient", **k0), dict(type="symmetryPlane", **k0)
    bcs = {"U": {"inlet": fv, "outlet": zg, "bottom": sym, "top": sym, "obstacle": sym, "defaultFaces": zg},
           "T": {"inlet": fv, "outlet": zg, "bottom": sym, "top": sym, "obstacle": zg, "defaultFaces": zg},
           "p": {"inlet": fv, "outlet": zg, "bottom": sym, "top": sym, "obstacle": zg, "defaultFaces": zg}}
    nin = mesh.boundary["inlet"]["nFaces"]
    Uin = np.zeros((nin, 3)); Uin[:, 0] = 3.
    bcvals = {("U", "inlet", "value"): Uin, ("T", "inlet", "value"): np.ones((nin, 1)), ("p", "inlet", "value"): np.ones((nin, 1))}
    spec = _spec(mesh, bcs, {"kind": "patch_pA", "patch": "obstacle"}, mu={"law": "constant", "value": 0.}, Cp=2.5)
    src = gaussian_source(mesh.cellCentres[:C], (0.58, 0.1, 0.), 1e-2, 2e3)
    case = Case(mesh, spec, conservative(U, T, p, Cp=2.5), src, bcvals, dt, dtype)
    case.primitive = (U, T, p)
    return case


def cylinder_shipped(dtype=np.float64, dt=2e-9):
    """BASELINE.json config 2 on the mesh the reference ships: cases/cylinder/constant/polyMesh/blockMeshDict (upper half of the
    cylinder, 10 blocks with arc edges, 46 250 cells; generated by adfvm_b200.blockmesh, patch table = the shipped `boundary` file)
    with the two z planes made cyclic (z1 / z2) as the case's create_mesh.sh does. Boundary conditions of cases/cylinder/0 where the
    reference has a class for them: no-slip cylinder (fixedValue U, zeroGradient T, p), CBC_UPT inflow on `left` (U0 = (33,0,0),
    T0 = 300, p0 = 102325) with the Lax-Friedrichs boundary solver, isothermal no-slip `up`; the outlet `right` is
    zeroGradient U, T + fixedValue p = 101325 (the shipped nonReflectingOutletPressure has no class in adFVM/BCs.py), `down` -
    the mirror plane of create_mesh.sh - a symmetryPlane. mu = 2.5e-5, drag objective and upstream source perturbation of
    templates/cylinder.py:9-19,66-78. Initial state: potential flow around the cylinder (the shipped uniform zero field has no
    dynamics in a few steps)."""
    from . import blockmesh
    mesh = build_mesh(blockmesh.block_mesh(**blockmesh.cylinder_dict(cyclic_span=True)))
    mesh.boundary["left"]["type"] = "characteristic"          # what CBC_UPT.__init__ does (BCs.py:142)
    mesh.boundary["down"]["type"] = "symmetryPlane"
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    r0, U0 = 0.5 * 2.5e-4, 33.
    r = np.linalg.norm(cc[:, :2], axis=1)
    th = np.arctan2(cc[:, 1], cc[:, 0])
    f = 1 - (r0 / r) ** 2
    U = np.stack([U0 * (1 - (r0 / r) ** 2 * np.cos(2 * th)) * f, -U0 * (r0 / r) ** 2 * np.sin(2 * th) * f, 0 * r], axis=1)
    T = np.full((C, 1), 300.)
    p = (101325 + 0.5 * 1.17 * (U0 ** 2 - (U ** 2).sum(axis=1))).reshape(-1, 1)
    k0 = {"keys": []}
    fv = {"type": "fixedValue", "keys": ["value"]}
    zg, sym, cyc, calc = dict(type="zeroGradient", **k0), dict(type="symmetryPlane", **k0), dict(type="cyclic", **k0), dict(type="calculated", **k0)
    bcs = {"U": {"down": sym, "right": zg, "up": fv, "left": calc, "cylinder": fv, "z1": cyc, "z2": cyc},
           "T": {"down": sym, "right": zg, "up": fv, "left": calc, "cylinder": zg, "z1": cyc, "z2": cyc},
           "p": {"down": sym, "right": fv, "up": zg, "left": {"type": "CBC_UPT", "keys": ["U0", "T0", "p0"]}, "cylinder": zg, "z1": cyc, "z2": cyc}}
    n = {k: mesh.boundary[k]["nFaces"] for k in ("up", "left", "right", "cylinder")}
    Ul = np.zeros((n["left"], 3)); Ul[:, 0] = U0
    bcvals = {("U", "up", "value"): np.zeros((n["up"], 3)), ("U", "cylinder", "value"): np.zeros((n["cylinder"], 3)),
              ("T", "up", "value"): np.full((n["up"], 1), 300.), ("p", "right", "value"): np.full((n["right"], 1), 101325.),
              ("p", "left", "U0"): Ul, ("p", "left", "T0"): np.full((n["left"], 1), 300.), ("p", "left", "p0"): np.full((n["left"], 1), 102325.)}
    spec = _spec(mesh, bcs, {"kind": "drag", "patch": "cylinder", "direction": 0}, mu={"law": "constant", "value": 2.5e-5},
                 briemann="eulerLaxFriedrichs")
    case = Case(mesh, spec, conservative(U, T, p), gaussian_source(cc, (-0.0005, 0., 0.), 1e-3, 2.5e6), bcvals, dt, dtype)
    case.primitive = (U, T, p)
    return case


def vane_cascade(nz=4, span=10.0, dtype=np.float64, dt=2e-8):
    """BASELINE.json config 4: one passage of the turbine-vane cascade of cases/vane_optim/foam/laminar/constant/polyMesh/
    blockMeshDict (16 blocks with spline edges, 10 000 cells per spanwise layer; adfvm_b200.blockmesh.vane_mesh), nz layers over
    `span` mm, with the set-up of templates/vane.py: total-pressure inlet (CBC_TOTAL_PT, pt = 175158 Pa), fixed-pressure outlet,
    isothermal no-slip blade surfaces `pressure` / `suction` (300 K), pitchwise and spanwise periodicity, Sutherland viscosity, the
    design objective of adFVM/objectives/vane.py (mass-flow averaged total-pressure loss over the plane x = 52.641 mm, cut by
    adfvm_b200.planecut as the reference's getPlane does; heat-transfer weights of getWeights), Gaussian source perturbation of
    templates/vane.py:24-35. The reference ships no field files for this case: the initial state is a smooth turning flow."""
    from . import blockmesh, planecut
    poly = blockmesh.vane_mesh(nz, span)
    mesh = build_mesh(poly)
    mesh.points, mesh.faces = poly.points, poly.faces
    mesh.boundary["inlet"]["type"] = "characteristic"         # what CBC_TOTAL_PT.__init__ does (BCs.py:167)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    s = np.clip((cc[:, 0] + 0.01) / 0.06, 0., 1.)              # 0 upstream of the blades, 1 downstream
    th = -np.deg2rad(55.) * s * s * (3 - 2 * s)
    speed = 60. + 90. * s
    U = np.stack([speed * np.cos(th), speed * np.sin(th), 0 * s], axis=1)
    T = (330. - 15. * s).reshape(-1, 1)
    p = (168000. - 30000. * s).reshape(-1, 1)
    k0 = {"keys": []}
    cyc, zg, calc = dict(type="cyclic", **k0), dict(type="zeroGradient", **k0), dict(type="calculated", **k0)
    fv = {"type": "fixedValue", "keys": ["value"]}
    per = {k: cyc for k in ("midplane1", "midplane2", "z1plane", "z2plane")}
    bcs = {"U": dict(per, inlet=calc, outlet=zg, pressure=fv, suction=fv),
           "T": dict(per, inlet=calc, outlet=zg, pressure=fv, suction=fv),
           "p": dict(per, inlet={"type": "CBC_TOTAL_PT", "keys": ["Tt", "pt"]}, outlet=fv, pressure=zg, suction=zg)}
    n = {k: mesh.boundary[k]["nFaces"] for k in ("inlet", "outlet", "pressure", "suction")}
    bcvals = {("U", "pressure", "value"): np.zeros((n["pressure"], 3)), ("U", "suction", "value"): np.zeros((n["suction"], 3)),
              ("T", "pressure", "value"): np.full((n["pressure"], 1), 300.), ("T", "suction", "value"): np.full((n["suction"], 1), 300.),
              ("p", "inlet", "Tt"): np.full((n["inlet"], 1), 340.), ("p", "inlet", "pt"): np.full((n["inlet"], 1), 175158.),
              ("p", "outlet", "value"): np.full((n["outlet"], 1), 138000.)}
    spec = _spec(mesh, bcs, {"kind": "plane_ptloss", "ptin": 175158., "normal": [1., 0., 0.], "scale": 0.4, "nExtra": 5})
    src = gaussian_source(cc, (0.05, -0.104, -0.5e-3 * span), 1e-1, 3e3)         # width of templates/vane.py:24-35, centred just upstream of the cut plane (the template's position needs thousands of steps to reach it)
    case = Case(mesh, spec, conservative(U, T, p), src, bcvals, dt, dtype)
    case.primitive = (U, T, p)
    cells, areas = planecut.intersect_plane(mesh, (0.052641, -0.1, -0.5e-3 * span), (1., 0., 0.))
    w = []
    for pid, (x0, y1) in (("pressure", (0.033757, 0.04692)), ("suction", (0.035241, 0.044337))):            # vane.py getWeights
        b = mesh.boundary[pid]
        fc = mesh.faceCentres[b["startFace"]:b["startFace"] + b["nFaces"]]
        w.append(np.ascontiguousarray(((fc[:, 0] >= x0) & (fc[:, 1] <= y1)).astype(case.dtype).reshape(-1, 1)))
    case.extra = [len(cells), np.ascontiguousarray(cells.reshape(-1, 1)), np.ascontiguousarray(areas.reshape(-1, 1), case.dtype)] + w
    case.extra_patches = ["pressure", "suction"]
    return case
