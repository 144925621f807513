#!/usr/bin/env python
"""Minimal driver for profiling: N resident primal + adjoint steps of the periodic box (no timing, no host copies).
Usage: python tools/run_step.py [--n 128] [--dtype f64] [--steps 2]   (wrap in ncu, see profiles/README.md)"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adfvm_b200 import cases, function  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--primal-only", action="store_true")
a = ap.parse_args()
dtype = np.float64 if a.dtype == "f64" else np.float32
case = cases.periodic_box(a.n, dtype)
f = function.PrimalFunction(case.spec, dtype)
fa = f.grad()
adj = [np.ascontiguousarray(np.ones_like(s) * w, dtype) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
f(*case.inputs(), replace_reusable=True, return_reusable=False)
if not a.primal_only:
    fa(*case.adjoint_inputs(case.state, adj), return_static=False)
for _ in range(a.steps):
    f.step_resident(case.dt)
    if not a.primal_only:
        fa.step_resident(case.dt, 1.0, chain=True)
f.sync()
print("done", f.launches)
