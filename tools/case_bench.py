#!/usr/bin/env python
"""Throughput on the shipped cases' own meshes (BASELINE.json configs 1-4: shockTube 500 cells, forwardStep 16 128 and cylinder
46 250 from their blockMeshDicts, the vane cascade with 4 spanwise layers = 40 000 cells): device-resident primal and adjoint steps, CUDA-event timing, one JSON line per
case and precision. These meshes are far too small to fill a B200 (launch-latency bound): the numbers are reported
for completeness next to the 368^3 headline of bench.py.
Usage: python tools/case_bench.py [--steps 200]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from adfvm_b200 import cases, function, hexmesh  # noqa: E402
from adfvm_b200.metrics import build_mesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=100)
a = ap.parse_args()


stream = torch.cuda.Stream()          # an explicit stream: whole steps replay as CUDA graphs (not possible on the legacy stream)
torch.cuda.set_stream(stream)
# the meshes of the shipped blockMeshDicts (adfvm_b200.blockmesh; parity on them: tests/test_anchors.py) - BASELINE.json configs 1-4
for name, make in (("shockTube_500", lambda d: cases.shock_tube(500, dtype=d)),
                   ("forwardStep_16128_shipped_mesh", lambda d: cases.forward_step_shipped(dtype=d)),
                   ("cylinder_46250_shipped_mesh", lambda d: cases.cylinder_shipped(dtype=d)),
                   ("vane_cascade_40000", lambda d: cases.vane_cascade(nz=4, dtype=d))):
    for dtype, tag in ((np.float64, "f64"), (np.float32, "f32")):
        case = make(dtype)
        C = case.mesh.nInternalCells
        f = function.PrimalFunction(case.spec, dtype, stream=stream.cuda_stream)
        fa = f.grad()
        adj = [np.ascontiguousarray(np.ones_like(s) * w, dtype) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
        f(*case.inputs(), replace_reusable=True, return_reusable=False)
        fa(*case.adjoint_inputs(case.state, adj), return_static=False)
        f(*case.inputs(), replace_reusable=True, return_reusable=False)
        for _ in range(10):
            f.step_resident(case.dt)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(a.steps):
            f.step_resident(case.dt)
        ev[1].record()
        for _ in range(a.steps):
            fa.step_resident(case.dt, 1.0, chain=True)
        ev[2].record()
        torch.cuda.synchronize()
        tp, ta = ev[0].elapsed_time(ev[1]) / a.steps, ev[1].elapsed_time(ev[2]) / a.steps
        dtc, obj = f.dtc_obj()
        print(json.dumps({"case": name, "dtype": tag, "cells": C, "steps": a.steps, "primal_ms_per_step": tp, "adjoint_ms_per_step": ta,
                          "primal_Mcell_updates_per_s": 3 * C / tp / 1e3, "adjoint_Mcell_updates_per_s": 3 * C / ta / 1e3,
                          "tile_stats": f.tile_stats(), "graph_replays": f.graph_replays, "finite": bool(np.isfinite(obj))}), flush=True)
