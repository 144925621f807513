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
ap.add_argument("--n", type=int, default=12)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = (a.n, a.n - 2, a.n - 4)


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
ids = decompose.global_cell_ids(N, rank, world)
gr = f.grad()(*case.adjoint_inputs(case.state, [np.ascontiguousarray(x[ids]) for x in adj]))
errs = [relerr(x, y[ids]) for x, y in zip(o[:3], out[:3])] + [relerr(x, y[ids]) for x, y in zip(o2[:3], out2[:3])]
errs.append(relerr(o[4], out[4]))
sc = [float(np.abs(s).max()) for s in g.state]
for grp in (slice(0, 3), slice(3, 6)):
    num = max(np.abs(x - y[ids]).max() * s for x, y, s in zip(gr[grp], grad[grp], sc))
    den = max(np.abs(y).max() * s for y, s in zip(grad[grp], sc))
    errs.append(num / den)
e = max(errs)
print("rank %d of %d maxerr %.3e launches %d" % (rank, world, e, f.launches), flush=True)
t = torch.tensor([e], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
dist.destroy_process_group()
sys.exit(0 if t.item() < 1e-10 else 1)
