"""Size-independent properties of the path (CPU simulator of the device code, 24^3 cells; the same checks hold at any
size): discrete conservation on a periodic box, linearity of the adjoint in its seeds, and the duality
<adjoint seeds, J v> = <J^T seeds, v> between the tangent of one step (central difference) and the adjoint."""
import numpy as np

from adfvm_b200 import cases, function


def test_conservation_on_periodic_box(hostsim):
    """without source terms the fluxes telescope: sum_c V_c W_c is unchanged by a step (SSPRK3 is a convex combination)"""
    case = cases.periodic_box(24, warp=0.02)
    zero = [np.zeros_like(s) for s in case.source]
    case.source = zero; case._static = None
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    out = f(*case.inputs(), replace_reusable=True, return_reusable=True)
    V = case.mesh.volumes
    for new, old in zip(out[:3], case.state):
        tot0, tot1 = (V * old).sum(axis=0), (V * new).sum(axis=0)
        assert np.all(np.abs(tot1 - tot0) <= 1e-11 * (V * np.abs(old)).sum(axis=0).max())


def test_adjoint_is_linear_in_its_seeds(hostsim):
    case = cases.walled_box((8, 6, 4))
    fa = function.PrimalFunction(case.spec, np.float64, lib=hostsim).grad()
    rng = np.random.RandomState(5)
    x = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    y = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    run = lambda seeds, obja: fa(*case.adjoint_inputs(case.state, seeds, obja=obja), zero_static=True)
    gx, gy = run(x, 1.0), run(y, 0.0)
    gz = run([2.0 * a - 3.0 * b for a, b in zip(x, y)], 2.0)
    sc = [float(np.abs(s).max()) for s in case.state]
    for grp in (slice(0, 3), slice(3, 6)):
        num = max(np.abs(c - (2.0 * a - 3.0 * b)).max() * s for a, b, c, s in zip(gx[grp], gy[grp], gz[grp], sc))
        den = max(np.abs(c).max() * s for c, s in zip(gz[grp], sc))
        assert num <= 1e-11 * den


def test_adjoint_tangent_duality(hostsim):
    """<a, dW'/dW v> (central difference of the primal step) = <(dW'/dW)^T a, v> (adjoint step, no objective)"""
    case = cases.walled_box((8, 6, 4))
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    rng = np.random.RandomState(11)
    a = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    v = [np.ascontiguousarray(rng.randn(*s.shape) * np.abs(s).max()) for s in case.state]
    eps = 1e-6
    step = lambda st: f(*case.inputs(st), replace_reusable=True, return_reusable=True)[:3]
    plus = step([s + eps * d for s, d in zip(case.state, v)])
    minus = step([s - eps * d for s, d in zip(case.state, v)])
    lhs = sum(float((x * (p - m)).sum()) for x, p, m in zip(a, plus, minus)) / (2 * eps)
    g = f.grad()(*case.adjoint_inputs(case.state, a, obja=0.0))
    rhs = sum(float((x * d).sum()) for x, d in zip(g[:3], v))
    assert abs(lhs - rhs) <= 1e-7 * abs(rhs)
