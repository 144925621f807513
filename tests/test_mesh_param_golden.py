"""parameters = 'mesh' against the UNMODIFIED reference: fixture box_walls_mesh (oracle/ref_harness/gen_golden.py) records the
`primal_grad` calls of apps/adjoint.py with the ten metric arrays as the parameter block (apps/adjoint.py:105-107; the
perturbed run dilates the points through the reference's own Mesh.getPointsPerturbation). Every one of the ten gradient
arrays is compared at its own scale: oracle 1e-12, C ABI (CPU simulator; the device in tests/test_gpu_parity.py) 1e-10."""
import numpy as np

from golden_util import Golden, group_relerr, state_scales
from adfvm_b200 import function
from oracle import adfvm_oracle as O

NAME = "box_walls_mesh"


def _each(r, out, tol):
    assert len(r) == len(out) == 13
    n = 0
    for a, b in zip(r[3:], out[3:]):
        b = np.asarray(b, np.float64).reshape(np.asarray(a).shape)
        assert np.abs(np.asarray(a) - b).max() <= tol * np.abs(b).max()
        n += np.abs(b).max() > 0
    return n


def test_oracle_mesh_gradients_match_reference():
    g = Golden(NAME)
    assert g.spec["parameters"] == "mesh"
    seen = 0
    for ci, nm, inp, opt, out in g.calls("adjoint", "primal_grad"):
        r = O.primal_grad(g.spec, inp)
        assert group_relerr(r[:3], out[:3], state_scales(inp)) < 1e-12
        seen = max(seen, _each(r, out, 1e-12))
    assert seen == 10


def replay(**kw):
    g = Golden(NAME)
    f = function.PrimalFunction(g.spec, np.float64, **kw).grad()
    for ci, nm, inp, opt, out in g.calls("adjoint", "primal_grad"):
        r = f(*inp, **opt)
        assert group_relerr(r[:3], out[:3], state_scales(inp)) < 1e-10
        _each(r, out, 1e-10)


def test_mesh_gradients_through_the_c_abi(hostsim):
    replay(lib=hostsim)
