#!/usr/bin/env python
"""Decomposition invariance on real GPUs (run under torchrun, one rank per GPU, NCCL halo inside the library):
every rank steps its block of a periodic box (2 primal steps + 1 adjoint step) and compares with the single-rank
run of the undecomposed mesh on its own GPU — the reference's criterion (tests/test_parallel.py:63-81).
Prints one line per rank `rank R maxerr E` and exits non-zero above 1e-10 (fp64).
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multigpu_check.py [--n 12]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from adfvm_b200 import decompose, function  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=48, help="cells per side of every rank's block (48: tiles that do not touch a processor "
                                                   "patch exist, so the overlapped early/late path really runs)")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = (a.n, a.n - 4, a.n - 8) if a.n >= 24 else (a.n, a.n - 2, a.n - 4)


def relerr(x, y):
    return float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))


g = decompose.global_box(N, world)
fs = function.PrimalFunction(g.spec, np.float64, device=local)
out = fs(*g.inputs(), replace_reusable=True)
out2 = fs(*g.inputs(list(out[:3])), replace_reusable=True)
rng = np.random.RandomState(7)
adj = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(g.state, (1.0, 1e-2, 1e-5))]
grad = fs.grad()(*g.adjoint_inputs(g.state, adj))

case = decompose.periodic_box_rank(N, rank, world)
f = function.PrimalFunction(case.spec, np.float64, device=local)
decompose.attach_comm(f, rank, world)
o = f(*case.inputs(), replace_reusable=True)
o2 = f(*case.inputs(list(o[:3])), replace_reusable=True)
early_tiles, all_tiles = f.tile_rounds()[2], f.tile_stats()[2]
if a.n >= 24:
    assert 0 < early_tiles < all_tiles, "the overlapped early/late split is not exercised (early %d of %d tiles)" % (early_tiles, all_tiles)
ids = decompose.global_cell_ids(N, rank, world)
gr = f.grad()(*case.adjoint_inputs(case.state, [np.ascontiguousarray(x[ids]) for x in adj]))
errs = [relerr(x, y[ids]) for x, y in zip(o[:3], out[:3])] + [relerr(x, y[ids]) for x, y in zip(o2[:3], out2[:3])]
errs.append(relerr(o[4], out[4]))
sc = [float(np.abs(s).max()) for s in g.state]
for grp in (slice(0, 3), slice(3, 6)):
    num = max(np.abs(x - y[ids]).max() * s for x, y, s in zip(gr[grp], grad[grp], sc))
    den = max(np.abs(y).max() * s for y, s in zip(grad[grp], sc))
    errs.append(num / den)
# ---- the vane design objective on the decomposed mesh: cut-plane cells as extraArgs, its mass flux all-reduced on the
# device (ncclAllReduce on the compute stream) between the two passes, forward and in the adjoint seeds
PLANE = {"kind": "plane_ptloss", "ptin": 175158., "normal": [1., 0., 0.], "scale": 0.4, "nExtra": 5}
cc = g.mesh.cellCentres[:g.mesh.nInternalCells]
xs = np.unique(np.round(cc[:, 0], 12))
sel = np.where(np.abs(cc[:, 0] - xs[len(xs) // 2]) < 1e-9)[0]
plane = [(int(c), 1e-3 * (1 + 0.1 * np.sin(c))) for c in sel]


def extras(ids_global):
    pos = {int(gid): i for i, gid in enumerate(ids_global)}
    mine = [(pos[c], a_) for c, a_ in plane if c in pos]
    return [len(mine), np.array([c for c, _ in mine], np.int32).reshape(-1, 1),
            np.array([a_ for _, a_ in mine], np.float64).reshape(-1, 1), np.zeros((1, 1)), np.zeros((1, 1))]


one = lambda v: np.array([[v]], np.float64)
gspec = dict(g.spec); gspec["objective"] = PLANE
fp = function.PrimalFunction(gspec, np.float64, device=local)
exg = extras(range(g.mesh.nInternalCells))
outp = fp(*(g.inputs() + exg), replace_reusable=True)
gradp = fp.grad()(*(g.inputs() + exg + adj + [one(0.), one(1.), one(0.)]))
cspec = dict(case.spec); cspec["objective"] = PLANE
fr = function.PrimalFunction(cspec, np.float64, device=local)
decompose.attach_comm(fr, rank, world)
exr = extras(ids)
op = fr(*(case.inputs() + exr), replace_reusable=True)
grp_ = fr.grad()(*(case.inputs() + exr + [np.ascontiguousarray(x[ids]) for x in adj] + [one(0.), one(1.), one(0.)]))
assert abs(outp[4][0, 0]) > 1e-6
errs.append(relerr(op[4], outp[4]))
for grp in (slice(0, 3), slice(3, 6)):
    num = max(np.abs(x - y[ids]).max() * s for x, y, s in zip(grp_[grp], gradp[grp], sc))
    den = max(np.abs(y).max() * s for y, s in zip(gradp[grp], sc))
    errs.append(num / den)
# ---- a walled (non-periodic) mesh decomposed like decomposePar: physical patches split between the ranks, processor
# patches from the cut faces (decompose.rank_cases)
from adfvm_b200 import cases  # noqa: E402
gw = cases.walled_box((8, 6, 4), warp=0.0)
fw = function.PrimalFunction(gw.spec, np.float64, device=local)
ow = fw(*gw.inputs(), replace_reusable=True)
adjw = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(gw.state, (1.0, 1e-2, 1e-5))]
gradw = fw.grad()(*gw.adjoint_inputs(gw.state, adjw))
cw, idw = decompose.rank_cases(gw, world)[rank]
fwr = function.PrimalFunction(cw.spec, np.float64, device=local)
decompose.attach_comm(fwr, rank, world)
owr = fwr(*cw.inputs(), replace_reusable=True)
gwr = fwr.grad()(*cw.adjoint_inputs(cw.state, [np.ascontiguousarray(x[idw]) for x in adjw]))
errs += [relerr(x, y[idw]) for x, y in zip(owr[:3], ow[:3])] + [relerr(owr[4], ow[4])]
scw = [float(np.abs(s).max()) for s in gw.state]
for grp in (slice(0, 3), slice(3, 6)):
    num = max(np.abs(x - y[idw]).max() * s for x, y, s in zip(gwr[grp], gradw[grp], scw))
    den = max(np.abs(y).max() * s for y, s in zip(gradw[grp], scw))
    errs.append(num / den)
# ---- adjoint artificial viscosity on the decomposed walled mesh: all-reduced normalisation of M_2norm, its halo, the diffusion
# solve across processor patches with the dot products all-reduced on the device (the walled mesh: the reference does not couple
# cells across cyclic patches, so on periodic meshes its own result depends on the decomposition - tests/test_viscosity.py)
SCAL = 3e4
def _scaled(inp):
    inp = list(inp); inp[-1] = np.array([[SCAL]], np.float64); return inp
vis_ref = fw.grad().viscous("abarbanel")(*_scaled(gw.adjoint_inputs(gw.state, adjw)))
vis_r = fwr.grad().viscous("abarbanel")(*_scaled(cw.adjoint_inputs(cw.state, [np.ascontiguousarray(x[idw]) for x in adjw])))
num = max(np.abs(x - y[idw]).max() * s for x, y, s in zip(vis_r[:3], vis_ref[:3], scw))
den = max(np.abs(y).max() * s for y, s in zip(vis_ref[:3], scw))
errs.append(num / den)
assert max(np.abs(x - y).max() * s for x, y, s in zip(vis_ref[:3], gradw[:3], scw)) > 1e-3 * den      # the smoothing acted
# ---- BASELINE.json config 4: the vane cascade of cases/vane_optim (spline-edged multi-block mesh, 4 spanwise layers = 40 000 cells)
# with the reference's design objective, decomposed like decomposePar (decompose.rank_cases: plane cells and weights per rank)
gv = cases.vane_cascade(nz=4)
fv = function.PrimalFunction(gv.spec, np.float64, device=local)
ov = fv(*gv.inputs(), replace_reusable=True)
ov2 = fv(*gv.inputs(list(ov[:3])), replace_reusable=True)
adjv = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(gv.state, (1.0, 1e-2, 1e-5))]
gradv = fv.grad()(*gv.adjoint_inputs(gv.state, adjv))
cv, idv = decompose.rank_cases(gv, world)[rank]
fvr = function.PrimalFunction(cv.spec, np.float64, device=local)
decompose.attach_comm(fvr, rank, world)
ovr = fvr(*cv.inputs(), replace_reusable=True)
ovr2 = fvr(*cv.inputs(list(ovr[:3])), replace_reusable=True)
gvr = fvr.grad()(*cv.adjoint_inputs(cv.state, [np.ascontiguousarray(x[idv]) for x in adjv]))
assert abs(ov[4][0, 0]) > 1e-3
errs += [relerr(x, y[idv]) for x, y in zip(ovr[:3], ov[:3])] + [relerr(x, y[idv]) for x, y in zip(ovr2[:3], ov2[:3])] + [relerr(ovr[4], ov[4])]
scv = [float(np.abs(s).max()) for s in gv.state]
for grp in (slice(0, 3), slice(3, 6)):
    num = max(np.abs(x - y[idv]).max() * s for x, y, s in zip(gvr[grp], gradv[grp], scv))
    den = max(np.abs(y).max() * s for y, s in zip(gradv[grp], scv))
    errs.append(num / den)
e = max(errs)
print("rank %d of %d maxerr %.3e launches %d early_tiles %d of %d" % (rank, world, e, f.launches, early_tiles, all_tiles), flush=True)
t = torch.tensor([e], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
dist.destroy_process_group()
sys.exit(0 if t.item() < 1e-10 else 1)
