#!/usr/bin/env python
"""Per-kernel times of the adjoint artificial viscosity (fvm_viscosity.h) on the periodic box: one adjoint step resident,
then the smoothing with kernel timing on. Prints ms per launch and the HBM rate of the CG kernels by their algorithmic bytes.
Usage: python tools/visc_bench.py [--n 256] [--dtype f64] [--type abarbanel] [--scaling 300]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adfvm_b200 import cases, function  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--type", default="abarbanel")
ap.add_argument("--scaling", type=float, default=300.)
a = ap.parse_args()
dtype = np.float64 if a.dtype == "f64" else np.float32
s = np.dtype(dtype).itemsize
case = cases.periodic_box(a.n, dtype)
C = case.mesh.nInternalCells
f = function.PrimalFunction(case.spec, dtype)
fa = f.grad()
fv = fa.viscous(a.type)
rng = np.random.RandomState(3)
adj = [np.ascontiguousarray(rng.randn(*x.shape) * w, dtype) for x, w in zip(case.state, (1.0, 1e-2, 1e-5))]
f(*case.inputs(), replace_reusable=True, return_reusable=False)
fa(*case.adjoint_inputs(case.state, adj), return_static=False)
fa.step_resident(case.dt, 1.0, chain=True)
fv.viscous_resident(case.dt, a.scaling)          # warm-up (allocates the work arrays)
f.sync()
fa.step_resident(case.dt, 1.0, chain=True)
f.kernel_timing(True)
t0 = time.perf_counter()
fv.viscous_resident(case.dt, a.scaling)
f.sync()
wall = time.perf_counter() - t0
rep = f.kernel_report()
f.kernel_timing(False)
# algorithmic bytes per cell: spmv reads 5 p + 6 coef + 1 diag + 6 nbr ints, writes 5 q; update reads/writes x, r (10+10) + p, q (10) + diag
bytes_cell = {"visc_cg_spmv": (5 + 6 + 1 + 5) * s + 24, "visc_cg_update": 31 * s, "visc_cg_dir": 16 * s,
              "visc_eig": (2 + 12 + 24 + 1 + 1) * s + 24 + 6 * s, "visc_coef": (1 + 6 * 4 + 7) * s + 48}
out = {"n": a.n, "cells": C, "dtype": a.dtype, "type": a.type, "scaling": a.scaling, "iterations": fv.viscosity_iterations,
       "wall_ms": wall * 1e3, "kernels": {}}
for name, (cnt, ms) in rep.items():
    per = ms / cnt
    k = {"launches": cnt, "ms_per_launch": per}
    if name in bytes_cell:
        k["GBs"] = bytes_cell[name] * C / per / 1e6
    out["kernels"][name] = k
print(json.dumps(out))
