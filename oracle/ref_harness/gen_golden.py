"""TEST INFRASTRUCTURE ONLY — generate the golden fixtures under tests/golden/ by running the
UNMODIFIED reference (apps/problem.py orig, apps/problem.py perturb, apps/adjoint.py) on small hex
meshes produced by adfvm_b200.hexmesh. Needs /root/reference; the fixtures travel, this does not.

usage: python oracle/ref_harness/gen_golden.py [case ...] [--fp32]
Each case leaves tests/golden/<case>.{npz,json}: every compiled-function call the reference made
(positional inputs, options, outputs), the static problem description, and objective.txt lines.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from adfvm_b200 import hexmesh  # noqa: E402
from adfvm_b200.metrics import build_mesh  # noqa: E402
import foam_io  # noqa: E402

SCRATCH = os.environ.get("ADFVM_GOLDEN_SCRATCH", "/tmp/adfvm_golden")
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASEFILE_HEAD = '''import numpy as np
from adFVM import config
from adFVM.density import RCF
from adFVM.mesh import Mesh
from adpy import tensor

def _allreduce(x):
    inputs = (x,)
    outputs = tuple([tensor.Zeros(y.shape) for y in inputs])
    (x,) = tensor.ExternalFunctionOp('mpi_allreduce', inputs, outputs).outputs
    return x
'''

OBJ_CELL_TV = '''
def _objTV(U, T, p, volumes):
    return (T*volumes).sum()
def objective(fields, solver):
    U, T, p = fields
    mesh = solver.mesh.symMesh
    obj = tensor.Zeros((1,1))
    obj = tensor.Kernel(_objTV)(mesh.nInternalCells, (obj,))(U, T, p, mesh.volumes)
    return _allreduce(obj)
'''

OBJ_PATCH_PA = '''
def _objPA(U, T, p, *mesh, **options):
    mesh = Mesh.container(mesh)
    p0 = p.extract(mesh.neighbour)
    return (p0*mesh.areas).sum()
def objective(fields, solver):
    U, T, p = fields
    mesh = solver.mesh.symMesh
    patch = mesh.boundary['{patch}']
    startFace, nFaces = patch['startFace'], patch['nFaces']
    meshArgs = [x[startFace] for x in mesh.getTensor()]
    obj = tensor.Zeros((1,1))
    obj = tensor.Kernel(_objPA)(nFaces, (obj,))(U, T, p, *meshArgs)
    return _allreduce(obj)
'''

# the drag objective of reference templates/cylinder_test.py:9-36
OBJ_DRAG = '''
def _objDrag(U, T, p, *mesh, **options):
    solver = options['solver']
    mesh = Mesh.container(mesh)
    U0 = U.extract(mesh.neighbour)[0]
    U0i = U.extract(mesh.owner)[0]
    p0 = p.extract(mesh.neighbour)
    T0 = T.extract(mesh.neighbour)
    nx = mesh.normals[0]
    mungUx = solver.mu(T0)*(U0-U0i)/mesh.deltas
    drag = (p0*nx-mungUx)*mesh.areas
    return drag.sum()
def objective(fields, solver):
    U, T, p = fields
    mesh = solver.mesh.symMesh
    patch = mesh.boundary['{patch}']
    startFace, nFaces = patch['startFace'], patch['nFaces']
    meshArgs = [x[startFace] for x in mesh.getTensor()]
    obj = tensor.Zeros((1,1))
    obj = tensor.Kernel(_objDrag)(nFaces, (obj,))(U, T, p, *meshArgs, solver=solver)
    return _allreduce(obj)
'''

CASEFILE_TAIL = '''
primal = RCF('{case}/', objective=objective, fixedTimeStep=True{rcf_extra})
{after_primal}

def perturb(fields, mesh, t):
    x = mesh.cellCentres[:mesh.nInternalCells]
    mid = np.array({mid})
    G = ({amp}*np.exp(-{width}*np.linalg.norm(x-mid, axis=1, keepdims=1)**2)).astype(config.precision)
    rho = G
    rhoU = np.zeros((mesh.nInternalCells, 3), config.precision)
    rhoU[:, 0] += G.flatten()*100
    rhoE = (G*2e5).astype(config.precision)
    return rho, rhoU, rhoE

parameters = 'source'
{param_block}
nSteps = {nSteps}
writeInterval = {writeInterval}
startTime = 0.0
dt = {dt}
'''


def smooth_fields(cc, lo, hi):
    s3 = (cc - np.asarray(lo)) / (np.asarray(hi) - np.asarray(lo))
    s = np.sin(2 * np.pi * s3[:, 0]) * np.cos(2 * np.pi * s3[:, 1]) * np.sin(2 * np.pi * s3[:, 2])
    U = np.stack([100 + 10 * s, 50 - 5 * s, 20 + 2 * s], axis=1)
    T = (300 + 10 * s).reshape(-1, 1)
    p = (101325 + 1000 * s).reshape(-1, 1)
    return U, T, p


def case_box_cyclic():
    """3-D periodic box, smoothly warped (non-orthogonal), Sutherland viscosity, objective sum T*V."""
    lo, hi = (0., 0., 0.), (1., 1., 1.)
    poly = hexmesh.box_mesh((6, 5, 4), lo, hi, warp=hexmesh.sine_warp(0.03, lo, hi))
    m = build_mesh(poly)
    U, T, p = smooth_fields(m.cellCentres[:m.nInternalCells], lo, hi)
    bf = {k: {"type": "cyclic"} for k in poly.boundary}
    return dict(poly=poly, fields={"U": (U, bf), "T": (T, bf), "p": (p, bf)}, objective=OBJ_CELL_TV,
                obj_spec={"kind": "cell_TV"}, rcf_extra="", mid="[0.5,0.5,0.5]", amp="1e2", width="50", nSteps=4, writeInterval=2, dt=1e-6)


def case_tube():
    """Sod tube geometry (reference cases/shockTube/constant/polyMesh/blockMeshDict), 100 cells,
    tanh-smoothed IC (the sharp IC NaNs in the reference), inviscid."""
    lo, hi = (-5., -1., -1.), (5., 1., 1.)
    poly = hexmesh.box_mesh((100, 1, 1), lo, hi, patches=[
        ("sides", "patch", ["x+", "x-"], {}),
        ("empty", "empty", ["y-", "z+", "y+", "z-"], {})])
    m = build_mesh(poly)
    x = m.cellCentres[:m.nInternalCells, 0]
    sig = 0.5 * (1 - np.tanh(x / 0.3))
    pr = 1e4 + 9e4 * sig
    rho = 0.125 + 0.875 * sig
    R = 1004.5 - 1004.5 / 1.4
    T = (pr / (rho * R)).reshape(-1, 1)
    U = np.zeros((len(x), 3))
    bf = {"sides": {"type": "zeroGradient"}, "empty": {"type": "empty"}}
    return dict(poly=poly, fields={"U": (U, bf), "T": (T, bf), "p": (pr.reshape(-1, 1), bf)},
                objective=OBJ_PATCH_PA.format(patch="sides"),
                obj_spec={"kind": "patch_pA", "patch": "sides"}, rcf_extra=", mu=lambda T: 0.",
                mid="[-4.5,0.,0.]", amp="1e3", width="25", nSteps=6, writeInterval=3, dt=1e-5)


def case_box_walls():
    """Graded channel: total-pressure inlet (CBC_TOTAL_PT, characteristic flux), fixedValue-p outlet,
    symmetryPlane floor, no-slip isothermal lid (fixedValue U, T), cyclic span; constant viscosity;
    drag objective on the lid (reference templates/cylinder_test.py:9-36)."""
    lo, hi = (0., 0., 0.), (2., 1., 0.5)
    poly = hexmesh.box_mesh((8, 6, 3), lo, hi, grading=(1.0, 0.4, 1.0), patches=[
        ("inlet", "patch", ["x-"], {}),
        ("outlet", "patch", ["x+"], {}),
        ("floor", "symmetryPlane", ["y-"], {}),
        ("lid", "patch", ["y+"], {}),
        ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}),
        ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    m = build_mesh(poly)
    cc = m.cellCentres[:m.nInternalCells]
    s3 = (cc - np.asarray(lo)) / (np.asarray(hi) - np.asarray(lo))
    s = np.sin(2 * np.pi * s3[:, 0]) * np.cos(np.pi * s3[:, 1]) * np.cos(2 * np.pi * s3[:, 2])
    U = np.stack([60 * (1 - s3[:, 1] ** 2) + 3 * s, 2 * s, 1 * s], axis=1)
    T = (300 + 5 * s).reshape(-1, 1)
    p = (101325 + 500 * s).reshape(-1, 1)
    cyc = {"z1": {"type": "cyclic"}, "z2": {"type": "cyclic"}}
    nin = poly.boundary["inlet"]["nFaces"]
    nlid = poly.boundary["lid"]["nFaces"]
    ptv = 103000. + 100. * np.sin(np.arange(nin)).reshape(-1, 1)
    Ulid = np.zeros((nlid, 3)); Ulid[:, 0] = 1.0 + 0.1 * np.cos(np.arange(nlid))
    bU = dict(cyc, inlet={"type": "calculated"}, outlet={"type": "zeroGradient"},
              floor={"type": "symmetryPlane"}, lid={"type": "fixedValue", "value": "uniform (1.5 0 0)"})
    bT = dict(cyc, inlet={"type": "calculated"}, outlet={"type": "zeroGradient"},
              floor={"type": "symmetryPlane"}, lid={"type": "fixedValue", "value": "uniform 310"})
    bp = dict(cyc, inlet={"type": "CBC_TOTAL_PT", "Tt": "uniform 305", "pt": "uniform 103000", "value": "uniform 101325"},
              outlet={"type": "fixedValue", "value": "uniform 101000"},
              floor={"type": "symmetryPlane"}, lid={"type": "zeroGradient"})
    return dict(poly=poly, fields={"U": (U, bU), "T": (T, bT), "p": (p, bp)},
                objective=OBJ_DRAG.format(patch="lid"),
                obj_spec={"kind": "drag", "patch": "lid", "direction": 0}, rcf_extra=", mu=lambda T: 2.5e-5",
                mid="[1.0,0.5,0.25]", amp="1e2", width="20", nSteps=4, writeInterval=2, dt=2e-6)


def case_box_upt():
    """Warped box: CBC_UPT inlet with the Lax-Friedrichs boundary Riemann solver (as
    templates/cylinder_test.py:38-43), zeroGradient outlet, slip + empty walls, cyclic span,
    Sutherland viscosity, pressure-force objective on the outlet."""
    lo, hi = (0., 0., 0.), (1., 0.5, 0.5)
    poly = hexmesh.box_mesh((7, 4, 3), lo, hi, warp=hexmesh.sine_warp(0.02, lo, hi), patches=[
        ("left", "patch", ["x-"], {}),
        ("right", "patch", ["x+"], {}),
        ("down", "patch", ["y-"], {}),
        ("up", "empty", ["y+"], {}),
        ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}),
        ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    m = build_mesh(poly)
    U, T, p = smooth_fields(m.cellCentres[:m.nInternalCells], lo, hi)
    U[:, 0] -= 60.
    cyc = {"z1": {"type": "cyclic"}, "z2": {"type": "cyclic"}}
    nl = poly.boundary["left"]["nFaces"]
    U0 = np.zeros((nl, 3)); U0[:, 0] = 40 + np.sin(np.arange(nl)); U0[:, 1] = 50.
    bU = dict(cyc, left={"type": "calculated"}, right={"type": "zeroGradient"}, down={"type": "slip"}, up={"type": "empty"})
    bT = dict(cyc, left={"type": "calculated"}, right={"type": "zeroGradient"}, down={"type": "zeroGradient"}, up={"type": "empty"})
    bp = dict(cyc, left={"type": "CBC_UPT", "U0": "uniform (40 50 0)", "T0": "uniform 300", "p0": "uniform 101825",
                         "value": "uniform 101825"},
              right={"type": "zeroGradient"}, down={"type": "zeroGradient"}, up={"type": "empty"})
    return dict(poly=poly, fields={"U": (U, bU), "T": (T, bT), "p": (p, bp)},
                objective=OBJ_PATCH_PA.format(patch="right"), obj_spec={"kind": "patch_pA", "patch": "right"},
                rcf_extra=", boundaryRiemannSolver='eulerLaxFriedrichs'",
                mid="[0.5,0.25,0.25]", amp="1e2", width="60", nSteps=4, writeInterval=2, dt=1e-6)


def case_cyl2d():
    """2-D laminar cylinder in the spirit of reference templates/cylinder.py (config 2 of BASELINE.json): half
    annulus around a cylinder of radius 0.5 mm, geometric radial grading, one cell in the span with cyclic z1/z2 (as
    cases/cylinder/0/U), no-slip wall (fixedValue U, zeroGradient T and p), CBC_UPT far field with the
    Lax-Friedrichs boundary Riemann solver, symmetry planes on the axis, mu = 2.5e-5, drag objective on the
    cylinder (templates/cylinder.py:9-19), source perturbation just upstream of the cylinder (:62-73)."""
    r0, r1, nr, nt = 0.5e-3, 8e-3, 12, 16
    def warp(p):
        r = r0 * (r1 / r0) ** p[:, 0]
        th = np.pi * p[:, 1]
        return np.stack([r * np.cos(th), r * np.sin(th), p[:, 2]], axis=1)
    poly = hexmesh.box_mesh((nr, nt, 1), (0., 0., 0.), (1., 1., 2e-4), warp=warp, patches=[
        ("cylinder", "patch", ["x-"], {}),
        ("far", "patch", ["x+"], {}),
        ("axis", "symmetryPlane", ["y-", "y+"], {}),
        ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}),
        ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    m = build_mesh(poly)
    cc = m.cellCentres[:m.nInternalCells]
    r = np.linalg.norm(cc[:, :2], axis=1)
    # potential-flow-like start: U = U0 (1 - r0^2/r^2) blended to zero at the wall
    U0 = 33.
    th = np.arctan2(cc[:, 1], cc[:, 0])
    f = 1 - (r0 / r) ** 2
    U = np.stack([U0 * (1 - (r0 / r) ** 2 * np.cos(2 * th)) * f, -U0 * (r0 / r) ** 2 * np.sin(2 * th) * f, 0 * r], axis=1)
    T = (300 + 2 * np.exp(-((r - r0) / 1e-3) ** 2)).reshape(-1, 1)
    p = (101325 + 0.5 * 1.17 * (U0 ** 2 - (U ** 2).sum(axis=1))).reshape(-1, 1)
    cyc = {"z1": {"type": "cyclic"}, "z2": {"type": "cyclic"}}
    bU = dict(cyc, cylinder={"type": "fixedValue", "value": "uniform (0 0 0)"}, far={"type": "calculated"}, axis={"type": "symmetryPlane"})
    bT = dict(cyc, cylinder={"type": "zeroGradient"}, far={"type": "calculated"}, axis={"type": "symmetryPlane"})
    bp = dict(cyc, cylinder={"type": "zeroGradient"}, axis={"type": "symmetryPlane"},
              far={"type": "CBC_UPT", "U0": "uniform (33 0 0)", "T0": "uniform 300", "p0": "uniform 101325", "value": "uniform 101325"})
    return dict(poly=poly, fields={"U": (U, bU), "T": (T, bT), "p": (p, bp)},
                objective=OBJ_DRAG.format(patch="cylinder"), obj_spec={"kind": "drag", "patch": "cylinder", "direction": 0},
                rcf_extra=", mu=lambda T: 2.5e-5, boundaryRiemannSolver='eulerLaxFriedrichs'",
                mid="[-0.8e-3,0.1e-3,1e-4]", amp="1e-1", width="2.5e6", nSteps=4, writeInterval=2, dt=2e-9)



def case_step2d():
    """Mach-3 forward-facing step of reference templates/forwardStep.py / cases/forwardStep (config 3 of BASELINE.json)
    on a coarse mesh: box [0,3]x[0,1] minus the step x>0.6, y<0.2 (the three blocks of its blockMeshDict), one cell
    in the span with `empty` faces, non-dimensional gas Cp=2.5 (c=1 at T=1), inviscid (mu=0), fixedValue inlet,
    inletOutlet/zeroGradient outlet, symmetryPlane top and bottom, slip obstacle, pressure-force objective on the
    obstacle (templates/forwardStep.py:8-28), source perturbation at the inlet cells (:32-40 uses the same idea)."""
    nx, ny = 30, 10
    lo, hi = (0., 0., -0.05), (3., 1., 0.05)
    K, J, I = np.meshgrid(np.arange(1), np.arange(ny), np.arange(nx), indexing="ij")
    keep = ~((I >= 6) & (J < 2))
    poly = hexmesh.masked_box_mesh((nx, ny, 1), lo, hi, keep, [
        ("inlet", "patch", ["x-"], {}),
        ("outlet", "patch", ["x+"], {}),
        ("bottom", "symmetryPlane", ["y-"], {}),
        ("top", "symmetryPlane", ["y+"], {}),
        ("defaultFaces", "empty", ["z-", "z+"], {})], hole=("obstacle", "patch", {}))
    m = build_mesh(poly)
    cc = m.cellCentres[:m.nInternalCells]
    s = np.sin(2 * np.pi * cc[:, 0] / 3.) * np.cos(np.pi * cc[:, 1])
    U = np.stack([3. + 0.05 * s, 0.02 * s, 0 * s], axis=1)
    T = (1. + 0.01 * s).reshape(-1, 1)
    p = (1. + 0.02 * s).reshape(-1, 1)
    bU = {"inlet": {"type": "fixedValue", "value": "uniform (3 0 0)"}, "outlet": {"type": "inletOutlet", "inletValue": "uniform (3 0 0)", "value": "uniform (3 0 0)"},
          "bottom": {"type": "symmetryPlane"}, "top": {"type": "symmetryPlane"}, "obstacle": {"type": "slip"}, "defaultFaces": {"type": "empty"}}
    bT = {"inlet": {"type": "fixedValue", "value": "uniform 1"}, "outlet": {"type": "inletOutlet", "inletValue": "uniform 1", "value": "uniform 1"},
          "bottom": {"type": "symmetryPlane"}, "top": {"type": "symmetryPlane"}, "obstacle": {"type": "zeroGradient"}, "defaultFaces": {"type": "empty"}}
    bp = {"inlet": {"type": "fixedValue", "value": "uniform 1"}, "outlet": {"type": "zeroGradient"},
          "bottom": {"type": "symmetryPlane"}, "top": {"type": "symmetryPlane"}, "obstacle": {"type": "zeroGradient"}, "defaultFaces": {"type": "empty"}}
    return dict(poly=poly, fields={"U": (U, bU), "T": (T, bT), "p": (p, bp)},
                objective=OBJ_PATCH_PA.format(patch="obstacle"), obj_spec={"kind": "patch_pA", "patch": "obstacle"},
                rcf_extra=", Cp=2.5, mu=lambda T: 0., CFL=1.2", mid="[0.3,0.3,0.]", amp="1e-2", width="20", nSteps=4, writeInterval=2, dt=1e-3)



OBJ_VANE = '''
from adFVM.objectives.vane import objective, getWeights
'''

VANE_EXTRA = '''
# cut plane of the pressure-loss objective (adFVM/objectives/vane.py:25-38 finds it with the Cython intersectPlane,
# stubbed in this harness): the cells of the column x = xc and their cross-section areas, passed the same way
_m = primal.mesh
_cc = _m.cellCentres[:_m.nInternalCells]
_sel = np.where(np.abs(_cc[:, 0] - {xc}) < 1e-9)[0].astype(np.int32)
_area = _m.volumes[_sel] / {dx}
primal.extraArgs.append((tensor.IntegerScalar(), len(_sel)))
_n = primal.extraArgs[-1][0]
primal.extraArgs.append((tensor.StaticIntegerVariable((_n, 1)), _sel.reshape(-1, 1)))
primal.extraArgs.append((tensor.StaticVariable((_n, 1)), _area.reshape(-1, 1).astype(config.precision)))
getWeights(primal)
'''


def case_channel_vane():
    """Design objective of reference templates/vane.py (config 4 of BASELINE.json): adFVM/objectives/vane.py - mass-flow
    averaged total-pressure loss over a cut plane (cell list + areas passed as extraArgs) and wall heat transfer on the
    patches `pressure` / `suction` with coordinate-selected weights (coefficient b = 0 in the reference) - on a small
    graded channel: total-pressure inlet, fixedValue-p outlet, isothermal no-slip `pressure` and `suction` walls,
    cyclic span, Sutherland viscosity."""
    lo, hi = (0., 0., 0.), (0.1, 0.04, 0.01)
    nx = 10
    poly = hexmesh.box_mesh((nx, 6, 3), lo, hi, grading=(1.0, 0.5, 1.0), patches=[
        ("inlet", "patch", ["x-"], {}), ("outlet", "patch", ["x+"], {}),
        ("pressure", "patch", ["y-"], {}), ("suction", "patch", ["y+"], {}),
        ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}), ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    m = build_mesh(poly)
    cc = m.cellCentres[:m.nInternalCells]
    s3 = (cc - np.asarray(lo)) / (np.asarray(hi) - np.asarray(lo))
    s = np.sin(2 * np.pi * s3[:, 0]) * np.cos(np.pi * s3[:, 1]) * np.cos(2 * np.pi * s3[:, 2])
    U = np.stack([180 * (1 - (2 * s3[:, 1] - 1) ** 2) + 20 + 3 * s, 2 * s, 1 * s], axis=1)
    T = (320 + 5 * s).reshape(-1, 1)
    p = (140000 + 500 * s).reshape(-1, 1)
    cyc = {"z1": {"type": "cyclic"}, "z2": {"type": "cyclic"}}
    wallU = {"type": "fixedValue", "value": "uniform (0 0 0)"}
    wallT = {"type": "fixedValue", "value": "uniform 300"}
    bU = dict(cyc, inlet={"type": "calculated"}, outlet={"type": "zeroGradient"}, pressure=wallU, suction=wallU)
    bT = dict(cyc, inlet={"type": "calculated"}, outlet={"type": "zeroGradient"}, pressure=wallT, suction=wallT)
    bp = dict(cyc, inlet={"type": "CBC_TOTAL_PT", "Tt": "uniform 340", "pt": "uniform 175158", "value": "uniform 140000"},
              outlet={"type": "fixedValue", "value": "uniform 139000"}, pressure={"type": "zeroGradient"}, suction={"type": "zeroGradient"})
    dx = (hi[0] - lo[0]) / nx
    return dict(poly=poly, fields={"U": (U, bU), "T": (T, bT), "p": (p, bp)}, objective=OBJ_VANE,
                after_primal=VANE_EXTRA.format(xc=repr(lo[0] + 6.5 * dx), dx=repr(dx)),
                obj_spec={"kind": "plane_ptloss", "ptin": 175158., "normal": [1., 0., 0.], "scale": 0.4, "nExtra": 5},
                rcf_extra="", mid="[0.05,0.02,0.005]", amp="1e1", width="2e3", nSteps=4, writeInterval=2, dt=2e-8)



BC_PT_PARAM = '''
# sensitivity to a boundary-condition input instead of the source term (apps/adjoint.py:108-116, adFVM/solver.py:262-270):
# the total pressure of the CBC_TOTAL_PT inlet, perturbed face by face
def perturb(fields, mesh, t):
    n = mesh.boundary['inlet']['nFaces']
    return (50.*(1 + 0.5*np.cos(np.arange(n)))).reshape(-1, 1)

parameters = ('BCs', 'p', 'inlet', 'pt')
'''


def case_box_walls_bcpt():
    """box_walls with parameters = ('BCs', 'p', 'inlet', 'pt'): the adjoint returns the gradient with respect to the
    inlet total-pressure array (a14: BC parameter block) and the perturbed run perturbs that array."""
    c = case_box_walls()
    c["param_block"] = BC_PT_PARAM
    c["parameters"] = ["BCs", "p", "inlet", "pt"]
    return c



MESH_PARAM = '''
# mesh sensitivities (apps/adjoint.py:105-107, templates/cylinder.py:45-60): the parameter block is the ten metric arrays,
# the perturbation a uniform dilation of the points pushed through the reference's own metric build
def perturb(fields, mesh, t):
    if not hasattr(perturb, 'perturbation'):
        points = 1e-6*mesh.points
        perturb.perturbation = mesh.getPointsPerturbation(points)
    return perturb.perturbation

parameters = 'mesh'
'''


def case_box_walls_mesh():
    """box_walls with parameters = 'mesh': the adjoint returns the gradients with respect to the ten metric arrays"""
    c = case_box_walls()
    c["param_block"] = MESH_PARAM
    c["parameters"] = "mesh"
    return c


CASES = {"box_walls_mesh": case_box_walls_mesh, "box_walls_bcpt": case_box_walls_bcpt, "channel_vane": case_channel_vane, "step2d": case_step2d, "cyl2d": case_cyl2d, "box_cyclic": case_box_cyclic, "tube": case_tube, "box_walls": case_box_walls, "box_upt": case_box_upt}


# ---- BASELINE.md section 5 reproducibility anchors, at their own sizes. Too large to store call by call: the fixture keeps
# objective.txt, the per-step time series and a strided sample + norms of the final state / final adjoint fields.
def case_anchor_box48():
    """Anchor A: 48^3 periodic unit box, cell id = i + 48 (j + 48 k), 4 steps, dt = 1e-6, objective sum T*V"""
    lo, hi = (0., 0., 0.), (1., 1., 1.)
    poly = hexmesh.box_mesh((48, 48, 48), lo, hi)
    m = build_mesh(poly)
    U, T, p = smooth_fields(m.cellCentres[:m.nInternalCells], lo, hi)
    bf = {k: {"type": "cyclic"} for k in poly.boundary}
    return dict(poly=poly, fields={"U": (U, bf), "T": (T, bf), "p": (p, bf)}, objective=OBJ_CELL_TV,
                obj_spec={"kind": "cell_TV"}, rcf_extra="", mid="[0.5,0.5,0.5]", amp="1e2", width="50", nSteps=4, writeInterval=2, dt=1e-6,
                builder={"kind": "periodic_box", "n": [48, 48, 48]})


def case_anchor_tube500():
    """Anchor B: cases/shockTube geometry (500 x 1 x 1 cells on [-5,5] x [-1,1]^2), smoothed Sod initial condition
    sigma = (1 - tanh(x/0.06))/2, inviscid, objective sum_{sides} p_ghost * area, 20 steps, dt = 1e-5"""
    lo, hi = (-5., -1., -1.), (5., 1., 1.)
    poly = hexmesh.box_mesh((500, 1, 1), lo, hi, patches=[
        ("sides", "patch", ["x+", "x-"], {}),
        ("empty", "empty", ["y-", "z+", "y+", "z-"], {})])
    m = build_mesh(poly)
    x = m.cellCentres[:m.nInternalCells, 0]
    sig = 0.5 * (1 - np.tanh(x / 0.06))
    pr = 1e4 + 9e4 * sig
    rho = 0.125 + 0.875 * sig
    R = 1004.5 - 1004.5 / 1.4
    T = (pr / (rho * R)).reshape(-1, 1)
    U = np.zeros((len(x), 3))
    bf = {"sides": {"type": "zeroGradient"}, "empty": {"type": "empty"}}
    return dict(poly=poly, fields={"U": (U, bf), "T": (T, bf), "p": (pr.reshape(-1, 1), bf)},
                objective=OBJ_PATCH_PA.format(patch="sides"),
                obj_spec={"kind": "patch_pA", "patch": "sides"}, rcf_extra=", mu=lambda T: 0.",
                mid="[-4.5,0.,0.]", amp="1e3", width="25", nSteps=20, writeInterval=10, dt=1e-5,
                builder={"kind": "tube", "n": 500, "width": 0.06})


def case_anchor_forwardstep():
    """BASELINE.json config 3 as shipped: mesh of cases/forwardStep/constant/polyMesh/blockMeshDict (adfvm_b200.blockmesh; same patch
    table as the shipped `boundary`), the shipped uniform initial fields, the set-up of templates/forwardStep.py; Gaussian source
    perturbation just upstream of the step face (see adfvm_b200.cases.forward_step_shipped)"""
    from adfvm_b200 import blockmesh, cases
    poly = blockmesh.block_mesh(**blockmesh.forward_step_dict())
    c = cases.forward_step_shipped()
    U, T, p = c.primitive
    bU = {"inlet": {"type": "fixedValue", "value": "uniform (3 0 0)"}, "outlet": {"type": "inletOutlet", "inletValue": "uniform (3 0 0)", "value": "uniform (3 0 0)"},
          "bottom": {"type": "symmetryPlane"}, "top": {"type": "symmetryPlane"}, "obstacle": {"type": "slip"}, "defaultFaces": {"type": "empty"}}
    bT = {"inlet": {"type": "fixedValue", "value": "uniform 1"}, "outlet": {"type": "inletOutlet", "inletValue": "uniform 1", "value": "uniform 1"},
          "bottom": {"type": "symmetryPlane"}, "top": {"type": "symmetryPlane"}, "obstacle": {"type": "zeroGradient"}, "defaultFaces": {"type": "empty"}}
    bp = {"inlet": {"type": "fixedValue", "value": "uniform 1"}, "outlet": {"type": "zeroGradient"},
          "bottom": {"type": "symmetryPlane"}, "top": {"type": "symmetryPlane"}, "obstacle": {"type": "zeroGradient"}, "defaultFaces": {"type": "empty"}}
    return dict(poly=poly, fields={"U": (U, bU), "T": (T, bT), "p": (p, bp)},
                objective=OBJ_PATCH_PA.format(patch="obstacle"), obj_spec={"kind": "patch_pA", "patch": "obstacle"},
                rcf_extra=", Cp=2.5, mu=lambda T: 0., CFL=1.2", mid="[0.58,0.1,0.]", amp="1e-2", width="2e3", nSteps=20, writeInterval=10, dt=1e-4,
                builder={"kind": "forward_step_shipped"})


def case_anchor_cylinder():
    """BASELINE.json config 2 on the shipped mesh: cases/cylinder/constant/polyMesh/blockMeshDict (46 250 cells, arcs), z planes cyclic,
    boundary conditions of cases/cylinder/0 where the reference has classes for them (adfvm_b200.cases.cylinder_shipped), drag
    objective and upstream perturbation of templates/cylinder.py"""
    from adfvm_b200 import blockmesh, cases
    poly = blockmesh.block_mesh(**blockmesh.cylinder_dict(cyclic_span=True))
    poly.boundary["down"]["type"] = "symmetryPlane"
    c = cases.cylinder_shipped()
    U, T, p = c.primitive
    cyc = {"z1": {"type": "cyclic"}, "z2": {"type": "cyclic"}}
    bU = dict(cyc, down={"type": "symmetryPlane"}, right={"type": "zeroGradient"}, up={"type": "fixedValue", "value": "uniform (0 0 0)"},
              left={"type": "calculated"}, cylinder={"type": "fixedValue", "value": "uniform (0 0 0)"})
    bT = dict(cyc, down={"type": "symmetryPlane"}, right={"type": "zeroGradient"}, up={"type": "fixedValue", "value": "uniform 300"},
              left={"type": "calculated"}, cylinder={"type": "zeroGradient"})
    bp = dict(cyc, down={"type": "symmetryPlane"}, right={"type": "fixedValue", "value": "uniform 101325"}, up={"type": "zeroGradient"},
              left={"type": "CBC_UPT", "U0": "uniform (33 0 0)", "T0": "uniform 300", "p0": "uniform 102325", "value": "uniform 102325"},
              cylinder={"type": "zeroGradient"})
    return dict(poly=poly, fields={"U": (U, bU), "T": (T, bT), "p": (p, bp)},
                objective=OBJ_DRAG.format(patch="cylinder"), obj_spec={"kind": "drag", "patch": "cylinder", "direction": 0},
                rcf_extra=", mu=lambda T: 2.5e-5, boundaryRiemannSolver='eulerLaxFriedrichs'",
                mid="[-0.0005,0.,0.]", amp="1e-3", width="2.5e6", nSteps=10, writeInterval=5, dt=2e-9, builder={"kind": "cylinder_shipped"})


VANE_PLANE = '''
# templates/vane.py:18-20: the cut plane (adFVM/objectives/vane.py getPlane -> the reference's own Cython intersectPlane) and the
# heat-transfer weights become extraArgs of the step functions
from adFVM.objectives.vane import getPlane
getPlane(primal)
getWeights(primal)
'''


def case_anchor_vane():
    """BASELINE.json config 4: the vane cascade of cases/vane_optim/foam/laminar/constant/polyMesh/blockMeshDict, 4 spanwise layers
    (40 000 cells), set-up of templates/vane.py (adfvm_b200.cases.vane_cascade), design objective of adFVM/objectives/vane.py"""
    from adfvm_b200 import blockmesh, cases
    c = cases.vane_cascade(nz=4)
    poly = blockmesh.vane_mesh(4, 10.0)
    U, T, p = c.primitive
    cyc = {k: {"type": "cyclic"} for k in ("midplane1", "midplane2", "z1plane", "z2plane")}
    wallU = {"type": "fixedValue", "value": "uniform (0 0 0)"}
    wallT = {"type": "fixedValue", "value": "uniform 300"}
    bU = dict(cyc, inlet={"type": "calculated"}, outlet={"type": "zeroGradient"}, pressure=wallU, suction=wallU)
    bT = dict(cyc, inlet={"type": "calculated"}, outlet={"type": "zeroGradient"}, pressure=wallT, suction=wallT)
    bp = dict(cyc, inlet={"type": "CBC_TOTAL_PT", "Tt": "uniform 340", "pt": "uniform 175158", "value": "uniform 168000"},
              outlet={"type": "fixedValue", "value": "uniform 138000"}, pressure={"type": "zeroGradient"}, suction={"type": "zeroGradient"})
    return dict(poly=poly, fields={"U": (U, bU), "T": (T, bT), "p": (p, bp)}, objective=OBJ_VANE, after_primal=VANE_PLANE,
                obj_spec={"kind": "plane_ptloss", "ptin": 175158., "normal": [1., 0., 0.], "scale": 0.4, "nExtra": 5},
                rcf_extra="", mid="[0.05,-0.104,-0.005]", amp="1e-1", width="3e3", nSteps=4, writeInterval=2, dt=2e-8,
                builder={"kind": "vane_cascade", "nz": 4})


ANCHORS = {"anchor_vane": case_anchor_vane, "anchor_box48": case_anchor_box48, "anchor_tube500": case_anchor_tube500,
           "anchor_forwardstep": case_anchor_forwardstep, "anchor_cylinder": case_anchor_cylinder}
CASES.update(ANCHORS)          # write_case looks them up; the plain generator below skips them (see __main__)


def run(cmd, cwd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=cwd)


def write_case(name, tag):
    """case directory (mesh, fields, case file) of one golden case under SCRATCH; returns (dict, case dir, case file)"""
    c = CASES[name]()
    case = os.path.join(SCRATCH, tag)
    if os.path.exists(case):
        shutil.rmtree(case)
    os.makedirs(case)
    foam_io.write_polymesh(case, c["poly"])
    for fname, (internal, bf) in c["fields"].items():
        foam_io.write_field(case, "0", fname, internal, bf)
    casefile = os.path.join(case, "casefile.py")
    with open(casefile, "w") as f:
        f.write(CASEFILE_HEAD + c["objective"] + CASEFILE_TAIL.format(
            case=case, rcf_extra=c["rcf_extra"], after_primal=c.get("after_primal", ""), param_block=c.get("param_block", ""), mid=c["mid"], amp=c["amp"], width=c["width"],
            nSteps=c["nSteps"], writeInterval=c["writeInterval"], dt=c["dt"]))
    return c, case, casefile


def run_dropin(name, lib, runs=("orig", "perturb", "adjoint")):
    """Run the reference's own drivers (apps/problem.py, apps/adjoint.py, unmodified) on a golden case with their
    compiled `primal` / `primal_grad` functions replaced by adfvm_b200's Function objects over the native library
    `lib` (refshim drop-in mode). Returns the lines of objective.txt the reference wrote."""
    c, case, casefile = write_case(name, name + "_dropin")
    env = dict(os.environ, ADFVM_DROPIN_LIB=lib, ADFVM_DROPIN_OBJECTIVE=json.dumps(c["obj_spec"]))
    runner = os.path.join(HERE, "run_ref.py")
    py = sys.executable
    for r in runs:
        app = "adjoint" if r == "adjoint" else "problem"
        argv = [casefile, "-c"] + (["perturb"] if r == "perturb" else [])
        print("+ drop-in", r, flush=True)
        out = subprocess.run([py, runner, app, os.path.join(case, "rec_%s.npz" % r), "--"] + argv, cwd=case, env=env,
                             stdout=subprocess.PIPE, text=True, check=True).stdout
        served = [l for l in out.split("\n") if l.startswith("[dropin] served")]
        assert served and int(served[-1].split()[2]) > 0, "the drop-in functions were not called"
        print(served[-1], flush=True)
    with open(os.path.join(case, "objective.txt")) as f:
        return f.read().strip().split("\n")


def generate(name, fp32=False):
    tag = name + ("_fp32" if fp32 else "")
    c, case, casefile = write_case(name, tag)
    runner = os.path.join(HERE, "run_ref.py")
    flag = ["--fp32"] if fp32 else []
    py = sys.executable
    run([py, runner, "problem", os.path.join(case, "rec_orig.npz")] + flag + ["--", casefile, "-c"], case)
    run([py, runner, "problem", os.path.join(case, "rec_perturb.npz")] + flag + ["--", casefile, "-c", "perturb"], case)
    run([py, runner, "adjoint", os.path.join(case, "rec_adjoint.npz")] + flag + ["--", casefile, "-c"], case)

    # merge into one fixture
    out = {}
    meta = {"case": name, "fp32": fp32, "runs": {}}
    for runname in ("orig", "perturb", "adjoint"):
        z = np.load(os.path.join(case, "rec_%s.npz" % runname))
        with open(os.path.join(case, "rec_%s.json" % runname)) as f:
            j = json.load(f)
        meta["spec"] = j["spec"]
        meta["spec"]["objective"] = c["obj_spec"]
        if "parameters" in c:
            meta["spec"]["parameters"] = c["parameters"]
        meta["runs"][runname] = j["calls"]
        for k in z.files:
            out["%s__%s" % (runname, k)] = z[k]
    with open(os.path.join(case, "objective.txt")) as f:
        meta["objective_txt"] = f.read().strip().split("\n")
    meta["casefile"] = {k: c[k] for k in ("nSteps", "writeInterval", "dt", "mid", "amp", "width", "rcf_extra")}
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, tag + ".npz"), **out)
    with open(os.path.join(GOLDEN, tag + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", tag, meta["objective_txt"])


def _sample(arrs, stride):
    """strided rows + [sum, sum of squares, max |.|] per column of each array"""
    out = {}
    for k, a in arrs.items():
        a = np.asarray(a, np.float64).reshape(len(a), -1)
        out[k + "_sample"] = a[::stride].copy()
        out[k + "_norms"] = np.stack([a.sum(axis=0), (a * a).sum(axis=0), np.abs(a).max(axis=0)])
    return out


def generate_anchor(name, stride=97):
    """runs the unmodified reference like generate(); keeps objective.txt, time series and samples of the final fields"""
    c, case, casefile = write_case(name, name)
    runner = os.path.join(HERE, "run_ref.py")
    py = sys.executable
    run([py, runner, "problem", os.path.join(case, "rec_orig.npz"), "--", casefile, "-c"], case)
    run([py, runner, "problem", os.path.join(case, "rec_perturb.npz"), "--", casefile, "-c", "perturb"], case)
    run([py, runner, "adjoint", os.path.join(case, "rec_adjoint.npz"), "--", casefile, "-c"], case)
    out = {}
    meta = {"case": name, "stride": stride, "builder": c["builder"]}
    for runname, fname in (("orig", "primal"), ("perturb", "primal"), ("adjoint", "primal_grad")):
        z = np.load(os.path.join(case, "rec_%s.npz" % runname))
        with open(os.path.join(case, "rec_%s.json" % runname)) as f:
            j = json.load(f)
        meta["spec"] = j["spec"]
        meta["spec"]["objective"] = c["obj_spec"]
        last = max(ci for ci, cl in enumerate(j["calls"]) if cl["name"] == fname)
        meta[runname + "_calls"] = sum(1 for cl in j["calls"] if cl["name"] == fname)
        names = ("rho", "rhoU", "rhoE") if fname == "primal" else ("rhoa", "rhoUa", "rhoEa", "gS_rho", "gS_rhoU", "gS_rhoE")
        arrs = {"%s_%s" % (runname, n): z["c%d_o%d" % (last, oi)] for oi, n in enumerate(names)}
        out.update(_sample(arrs, stride))
        if runname == "orig":       # initial state (inputs 0-2 of the first call) and the source perturbation used by `perturb`
            first = min(ci for ci, cl in enumerate(j["calls"]) if cl["name"] == fname)
            out.update(_sample({"initial_%s" % n: z["c%d_i%d" % (first, ii)] for ii, n in enumerate(names)}, stride))
    with open(os.path.join(case, "objective.txt")) as f:
        meta["objective_txt"] = f.read().strip().split("\n")
    out["timeSeries_orig"] = np.loadtxt(os.path.join(case, "timeSeries.txt"), ndmin=1)
    out["timeSeries_perturb"] = np.loadtxt(os.path.join(case, "timeSeries_0.txt"), ndmin=1)
    out["sensTimeSeries"] = np.loadtxt(os.path.join(case, "sensTimeSeries.txt"), ndmin=1)
    meta["casefile"] = {k: c[k] for k in ("nSteps", "writeInterval", "dt", "mid", "amp", "width", "rcf_extra")}
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", name, meta["objective_txt"])


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    fp32 = "--fp32" in sys.argv
    for nm in (args or [n for n in CASES if n not in ANCHORS]):
        if nm in ANCHORS:
            generate_anchor(nm)
        else:
            generate(nm, fp32=fp32)
