"""Checkpoint blocks resident on the device: the two loops of the reference's drivers that round-trip
every state through numpy - `Solver.run(mode='forward')` (adFVM/solver.py:296-382: one `primal` call per step, every new state
appended to the host-side `solutions` list) and the step loop of `Adjoint.run` (apps/adjoint.py:217-291: one `primal_grad` call
per step with the stored state and the adjoint fields going up, the new adjoint fields coming down) - restated over
`PrimalFunction.run_block` / `AdjointFunction.run_block`, which keep the block's states, the adjoint fields and the static
gradient accumulator in HBM. The host sees one state per checkpoint (what the reference writes to disk every writeInterval steps)
and, per block, the adjoint fields and the accumulated parameter gradient.

A maintainer would call these from `Solver.run` / `Adjoint.run` in place of the per-step loops when no per-step report is wanted;
`tests/test_anchors.py` drives them against the reference's recorded anchor runs (objective, sensitivity, final fields)."""
import numpy as np


def forward_blocks(f, inputs, nSteps, writeInterval, dt):
    """`apps/problem.py` orig run: nSteps steps from the state in `inputs` (a `primal` positional list); returns the per-step
    objectives and the states at every multiple of writeInterval (index 0 = the initial state). dt: scalar or per-step sequence."""
    dts = [float(dt)] * nSteps if np.isscalar(dt) else [float(x) for x in dt]
    f.set_state(*inputs)
    series, checkpoints = [], [list(inputs[:3])]
    for b in range(nSteps // writeInterval):
        _, obj = f.run_block(dts[b * writeInterval:(b + 1) * writeInterval])
        series += [float(x) for x in obj]
        checkpoints.append(list(f.state()))              # fresh host arrays (page-locked), owned by the caller
    return series, checkpoints


def adjoint_blocks(f, fa, make_inputs, checkpoints, nSteps, writeInterval, dt, adjoint0, perturbation, obja=1.0,
                   viscous=None, scaling=0.0):
    """`apps/adjoint.py` Adjoint.run: checkpoints backwards, each block recomputed forward on the device and swept in reverse.
    make_inputs(state) -> `primal` positional list for that state; adjoint0: the adjoint fields of the final state;
    perturbation: arrays shaped like the parameter gradient (source terms), or None to skip the sensitivity sum. viscous: an `AdjointFunction.viscous(type)` object -
    the smoothing is then applied after every step (viscousInterval = 1). Returns (sum of sensitivities / nSteps, adjoint fields)."""
    dts = [float(dt)] * nSteps if np.isscalar(dt) else [float(x) for x in dt]
    total, out = 0.0, None
    for checkpoint in range(nSteps // writeInterval):
        k = nSteps // writeInterval - 1 - checkpoint
        blk = dts[k * writeInterval:(k + 1) * writeInterval]
        f.set_state(*make_inputs(checkpoints[k]))       # (also loads the static inputs on a fresh context)
        if checkpoint == 0:
            fa.set_fields(*adjoint0)
        f.run_block(blk)
        (viscous or fa).run_block(blk, obja, scaling)
        out = fa.fields(return_static=True, zero_static=True)
        if perturbation is not None:                    # cmesh.computeSensitivity (host work of the driver, apps/adjoint.py:341)
            total += sum(float((np.asarray(g, np.float64) * np.asarray(p, np.float64)).sum()) for g, p in zip(out[3:6], perturbation))
    return total / nSteps, list(out[:3])
