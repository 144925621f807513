"""Drop-in at the reference's own boundary: the UNMODIFIED reference drivers (apps/problem.py orig + perturb,
apps/adjoint.py) run a golden case with `solver.map` / `adjoint.map` served by adfvm_b200's PrimalFunction /
AdjointFunction (spec read off the reference's RCF object by spec_from_solver), and the
objective.txt they write must equal the one the reference wrote with its own compiled functions
(tests/golden/<case>.json) to 1e-9 - primal objective, finite-difference sensitivity and adjoint sensitivity.

Needs /root/reference (build container only; skipped elsewhere). The native library is the CPU simulator of the
device code here (no GPU in the build container); on a GPU box set ADFVM_DROPIN_USE_CUDA=1 to run the CUDA library."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ADFVM_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")


@pytest.mark.parametrize("name", ["box_walls"])
def test_reference_drivers_with_b200_functions(name, hostsim):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_harness"))
    import gen_golden
    from adfvm_b200 import _lib
    lib = _lib.DEFAULT_LIB if os.environ.get("ADFVM_DROPIN_USE_CUDA") else hostsim.path
    lines = gen_golden.run_dropin(name, lib)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))["objective_txt"]
    assert len(lines) == len(gold) == 3
    for got, ref in zip(lines, gold):
        g, r = got.split(), ref.split()
        assert g[0] == r[0]                                  # orig / perturb / adjoint
        # perturb = J(perturbed) - J(orig) cancels four digits: compare relative to the size of J
        scale = abs(float(gold[0].split()[3])) if g[0] == "perturb" else abs(float(r[3]))
        assert abs(float(g[3]) - float(r[3])) <= 1e-9 * scale, (got, ref)
    # the field files the reference wrote (final state of the `orig` run - the perturbed run does not write fields - and
    # the adjoint fields at the start time) hold what our
    # functions computed: read them back with the package's OpenFOAM reader and compare with the recorded stock run
    from adfvm_b200 import foam_io
    from golden_util import Golden, relerr
    case = os.path.join(gen_golden.SCRATCH, name + "_dropin")
    poly = foam_io.read_polymesh(case)
    n = int(poly.owner.max()) + 1
    gold_calls = Golden(name)
    last = [out for _, _, _, _, out in gold_calls.calls("orig", "primal") if out[0] is not None][-1]
    times = sorted((d for d in os.listdir(case) if d.replace(".", "").isdigit() and d != "0"), key=float)
    for fname, ref in zip(("rho", "rhoU", "rhoE"), last[:3]):
        got, _ = foam_io.read_field(case, times[-1], fname, n, poly.boundary)
        assert relerr(got, ref) < 1e-10, fname
    for fname in ("rhoa", "rhoUa", "rhoEa"):
        got, _ = foam_io.read_field(case, "0", fname, n, poly.boundary)
        assert np.all(np.isfinite(got)) and np.abs(got).max() > 0
