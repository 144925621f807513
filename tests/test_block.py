"""Device-side checkpoint block: PrimalFunction.run_block + AdjointFunction.run_block must
reproduce, bit for bit, the reference's way of doing it - every state of the block returned to the host by the primal
calls and fed back one by one to the adjoint calls (adFVM/solver.py:376-382, apps/adjoint.py:217-291) - on the CPU
simulator here and on the device in tests/test_gpu_parity.py."""
import numpy as np

from adfvm_b200 import cases, function


def block_vs_stepwise(lib, stream=None):
    case = cases.walled_box((8, 6, 4))
    dts = [case.dt, 0.8 * case.dt, 1.1 * case.dt, case.dt]
    adj = [np.ascontiguousarray(np.ones_like(s) * w) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    kw = {"lib": lib} if stream is None else {"stream": stream}
    # reference way: host round trips
    f = function.PrimalFunction(case.spec, np.float64, **kw)
    states, objs, state = [case.state], [], case.state
    for dt in dts:
        out = f(*case.inputs(state, dt), replace_reusable=True, return_reusable=True)
        state = [np.array(o) for o in out[:3]]
        states.append(state); objs.append(float(out[4][0, 0]))
    fa = f.grad()
    a = adj
    for k in reversed(range(len(dts))):
        g = fa(*case.adjoint_inputs(states[k], a, dt=dts[k]), return_static=True, zero_static=False)
        a = [np.array(x) for x in g[:3]]
    grads = [np.array(x) for x in g[3:6]]
    # block way: everything resident
    f2b = function.PrimalFunction(case.spec, np.float64, **kw)
    f2b.set_state(*case.inputs())
    dtc, obj = f2b.run_block(dts)
    fa2 = f2b.grad()
    fa2.set_fields(*adj)
    fa2.run_block(dts)
    res = fa2.fields(return_static=True)
    assert np.array_equal(obj, np.array(objs))
    for x, y in zip(res[:3], a):
        assert np.array_equal(x, y)
    for x, y in zip(res[3:6], grads):
        assert np.array_equal(x, y)
    # the resident primal state is back at the block's first state
    out = f2b(*case.inputs(dt=dts[0]), replace_reusable=False, return_reusable=True)
    for x, y in zip(out[:3], states[1]):
        assert np.array_equal(x, y)


def test_block_equals_host_round_trips(hostsim):
    block_vs_stepwise(hostsim)


def test_state_cache_opt_in(hostsim):
    """PrimalFunction(state_cache=n): returned states are read-only and have device copies; the very same array objects passed
    back (the `solutions` list of the reference's drivers) are not uploaded again; anything else is. Results are bitwise those of
    the default path."""
    import pytest
    case = cases.walled_box((8, 6, 4))
    adj = [np.ascontiguousarray(np.ones_like(s) * w) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]

    def run(cache):
        f = function.PrimalFunction(case.spec, np.float64, lib=hostsim, state_cache=cache)
        states, state = [case.state], case.state
        for k in range(3):
            out = f(*case.inputs(state), replace_reusable=(k == 0), return_reusable=True)
            state = list(out[:3]); states.append(state)
        fa, a, res = f.grad(), adj, []
        for k in reversed(range(3)):
            g = fa(*case.adjoint_inputs(states[k], a), return_static=True, zero_static=True)
            a = [np.array(x) for x in g[:3]]; res.append([np.array(x) for x in g])
        return f, states, res
    f0, s0, r0 = run(0)
    f1, s1, r1 = run(3)
    assert f0.state_cache_hits == 0 and s0[1][0].flags.writeable
    assert f1.state_cache_hits == 2                      # states[2] and states[1]; states[0] are the caller's own arrays
    for a, b in zip(r0, r1):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert not s1[1][0].flags.writeable
    with pytest.raises(ValueError):
        s1[2][1][0, 0] = 1.0                             # read-only
    # a state made writeable again and modified is uploaded (and the modification is seen)
    fa = f1.grad()
    st = s1[2]
    g_ref = fa(*case.adjoint_inputs(st, adj))
    hits = f1.state_cache_hits
    for x in st:
        x.flags.writeable = True
    st[0][0, 0] *= 1.01
    g_mod = fa(*case.adjoint_inputs(st, adj))
    assert f1.state_cache_hits == hits                   # no hit: uploaded
    assert not np.array_equal(g_mod[0], g_ref[0])
    # equal content in other array objects: uploaded as well
    g_copy = fa(*case.adjoint_inputs([np.array(x) for x in s1[1]], adj))
    assert f1.state_cache_hits == hits
    g_same = fa(*case.adjoint_inputs(s1[1], adj))
    assert f1.state_cache_hits == hits + 1
    for x, y in zip(g_copy[:3], g_same[:3]):             # (the static gradient accumulators keep summing across these calls)
        assert np.array_equal(x, y)
    # replace_reusable with cached arrays: the primal continues from the copy
    o_a = f1(*case.inputs(s1[1]), replace_reusable=True, return_reusable=True)
    o_b = f0(*case.inputs([np.array(x) for x in s0[1]]), replace_reusable=True, return_reusable=True)
    for x, y in zip(o_a[:3], o_b[:3]):
        assert np.array_equal(x, y)
