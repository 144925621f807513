"""Host-side tile plan (fvm_tiles.h) through the C ABI of the CPU simulator: structured blocks must decompose into
exact 4x4x8 tiles / 4x4x2 sub-tiles whatever the block size (4.0 flux evaluations per cell, 4 full rounds per sub-tile,
160 halo slots), including sizes that a plain halving of the tile count cuts raggedly (3*32, 5*32 planes) and the
sampled lattice detection used on large meshes."""
import os

import numpy as np
import pytest

from adfvm_b200 import cases, function


def _plan(hostsim, n):
    case = cases.periodic_box(n)
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    f.c.load_static(f.c.parse(case.inputs()))
    return f.tile_stats(), f.tile_rounds(), f.tile_halo_stats()


@pytest.mark.parametrize("n", [(16, 16, 16), (24, 24, 24), (40, 24, 8), (12, 20, 24)])
def test_structured_blocks_tile_regularly(hostsim, n):
    (evals, max_rounds, ntiles, T), (rounds, subtiles, early), (max_halo, variant, _) = _plan(hostsim, n)
    assert T == 128 and variant == 0
    assert evals == 4.0 and max_rounds == 4 and rounds == 4 * subtiles
    assert max_halo == 160
    assert early == ntiles                      # no processor patches: every tile is "early"


def test_sampled_lattice_detection(hostsim, monkeypatch):
    """large meshes detect their lattice from a scattered sample of the cells (a regular stride aliases with the
    lattice and misses planes); force the sampled path on a small mesh"""
    monkeypatch.setenv("ADFVM_LATTICE_SAMPLE", "3000")
    (evals, max_rounds, ntiles, T), (rounds, subtiles, _), (max_halo, variant, _) = _plan(hostsim, (24, 24, 24))
    assert evals == 4.0 and max_rounds == 4 and rounds == 4 * subtiles and max_halo == 160


def test_irregular_sizes_still_work(hostsim):
    """extents that are not multiples of 4: ragged tiles, more rounds, same results (checked by the golden tests)"""
    (evals, max_rounds, ntiles, T), (rounds, subtiles, _), _ = _plan(hostsim, (10, 9, 7))
    assert 3.0 < evals < 6.0 and max_rounds >= 4


def test_plan_independent_of_planner_threads(hostsim):
    """The tile planner runs on host threads (ADFVM_PLAN_THREADS, default min(cores, 16)): lattice detection, bisection and the
    per-chunk entry / schedule build must give the IDENTICAL plan for any thread count - checked through everything that
    depends on it: the outputs of a primal + adjoint step (bitwise), on a lattice mesh large enough for the threaded bisection
    and on a warped (non-lattice) mesh."""
    import hashlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
from adfvm_b200 import _lib, function, cases
lib = _lib.Lib(%r)
h = hashlib.sha256()
for case in (cases.periodic_box((64, 64, 64), np.float64), cases.periodic_box((40, 36, 28), np.float64, warp=0.03)):
    f = function.PrimalFunction(case.spec, np.float64, lib=lib)
    out = f(*case.inputs(), replace_reusable=True)
    adj = [np.ones_like(a) * w for a, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    g = f.grad()(*case.adjoint_inputs(case.state, adj))
    for a in list(out) + list(g):
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(np.asarray(f.tile_stats(), dtype=np.float64).tobytes())
print(h.hexdigest())
''' % (root, hostsim.path)
    digests = set()
    for nt in ("1", "5"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, ADFVM_PLAN_THREADS=nt), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        digests.add(r.stdout.strip().split("\n")[-1])
    assert len(digests) == 1, digests
