"""The shipped-case-sized builders of adfvm_b200.cases (cylinder2d: cases/cylinder at C = 46 250, forward_step:
cases/forwardStep at C = 16 128) are the golden fixtures `cyl2d` / `step2d` at another resolution: at the fixtures'
resolution they must reproduce, input by input, what the unmodified reference passed to its compiled `primal`
(mesh metrics from its cmesh build, state, BC arrays, patch table, BC classes)."""
import numpy as np
import pytest

from golden_util import Golden, relerr
from adfvm_b200 import cases


@pytest.mark.parametrize("name,make", [("cyl2d", lambda: cases.cylinder2d(12, 16, dt=2e-9)),
                                       ("step2d", lambda: cases.forward_step(30, 10, dt=1e-3))])
def test_builder_reproduces_golden_inputs(name, make):
    g = Golden(name)
    case = make()
    _, _, ref_in, _, _ = next(iter(g.calls("perturb", "primal")))        # the perturb run carries the source terms
    mine = case.inputs()
    assert len(mine) == len(ref_in)
    for k, (a, b) in enumerate(zip(mine, ref_in)):
        if isinstance(b, np.ndarray):
            assert a.shape == b.shape, k
            if b.dtype.kind == "f":
                assert relerr(a, b) < 1e-10, (k, relerr(a, b))
            else:
                assert np.array_equal(a, b), k
        else:
            assert int(a) == int(b), k
    for key in ("Cp", "gamma", "Pr", "mu", "riemannSolver", "boundaryRiemannSolver", "sortedPatches", "objective", "BCs"):
        assert g.spec[key] == case.spec[key], key
    for p, q in zip(g.spec["patches"], case.spec["patches"]):
        assert all(p[k] == q[k] for k in ("name", "type", "startFace", "nFaces", "cellStartFace")), (p, q)


def test_shipped_case_sizes():
    """cell / face counts of the reference's meshes (forwardStep C=16128, Fi=31936, F=64832)"""
    m = cases.forward_step().mesh
    assert (m.nInternalCells, m.nInternalFaces, m.nFaces) == (16128, 31936, 64832)
    for pid, n in (("inlet", 80), ("outlet", 64), ("bottom", 48), ("top", 240), ("obstacle", 208), ("defaultFaces", 32256)):
        assert m.boundary[pid]["nFaces"] == n
    assert cases.cylinder2d().mesh.nInternalCells == 46250
