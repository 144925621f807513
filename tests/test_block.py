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
