"""Pins the oracle (oracle/adfvm_oracle.py) against outputs of the UNMODIFIED reference recorded in
tests/golden (generator: oracle/ref_harness/gen_golden.py): every `primal` call (new state, dtc, objective)
and every `primal_grad` call (adjoint fields, source-term gradients). Tolerance 1e-12 relative (fp64;
the two differ only by summation order / pow vs sqrt round-off)."""
import numpy as np
import pytest

from golden_util import Golden, available, relerr, group_relerr, state_scales
from oracle import adfvm_oracle as O

CASES = [c for c in available() if not c.endswith("_fp32")]
TOL = 1e-12
# the drag on the half cylinder of `cyl2d` is a sum of pressure forces that cancel to 1e-4 of their size: the objective
# is compared at the north-star tolerance there
OBJ_TOL = {"cyl2d": 1e-10}
ADJ_TOL = {}


@pytest.mark.parametrize("name", CASES)
def test_oracle_primal_matches_reference(name):
    g = Golden(name)
    n = 0
    for run in ("orig", "perturb"):
        for ci, nm, inp, opt, out in g.calls(run, "primal"):
            r = O.primal(g.spec, inp)
            for a, b in zip(r[:3], out[:3]):
                assert relerr(a, b) < TOL
            assert relerr(r[3], out[3]) < TOL and relerr(r[4], out[4]) < OBJ_TOL.get(name, TOL)
            n += 1
    assert n >= 4


@pytest.mark.parametrize("name", CASES)
def test_oracle_adjoint_matches_reference(name):
    g = Golden(name)
    n = 0
    for ci, nm, inp, opt, out in g.calls("adjoint", "primal_grad"):
        r = O.primal_grad(g.spec, inp)
        sc = state_scales(inp)
        assert group_relerr(r[:3], out[:3], sc) < ADJ_TOL.get(name, TOL)
        assert group_relerr(r[3:6], out[3:6], sc) < ADJ_TOL.get(name, TOL)
        n += 1
    assert n >= 4


def test_objective_txt_anchor():
    """objective.txt of the recorded runs: adjoint sensitivity vs finite difference, the reference's own
    end-to-end criterion (tests/test_adjoint.py:33, 1e-3)."""
    for name in CASES:
        lines = Golden(name).meta["objective_txt"]
        fd = float(lines[1].split()[3]); adj = float(lines[2].split()[3])
        assert abs(fd - adj) / abs(fd) < 1e-3, (name, fd, adj)
