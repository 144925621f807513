#!/usr/bin/env python
"""Host-side set-up cost of one context (adfvm_set_mesh: validation, tile plan, uploads) with its phases (ADFVM_TILE_TIMING=1).
Usage: ADFVM_TILE_TIMING=1 python tools/setup_time.py [--n 256]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adfvm_b200 import cases, function  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
a = ap.parse_args()
t0 = time.time(); case = cases.periodic_box(a.n, np.float64); t1 = time.time()
print("case arrays (numpy, bench infrastructure) %.1f s" % (t1 - t0), flush=True)
f = function.PrimalFunction(case.spec, np.float64)
t1 = time.time(); f.set_state(*case.inputs()); t2 = time.time()
print("set_state: static upload + tile plan %.1f s for %d cells" % (t2 - t1, case.mesh.nInternalCells), flush=True)
