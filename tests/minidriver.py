"""Test helper: the time loops of the reference's drivers, restated just enough to reproduce `objective.txt`.

  run_primal   = apps/problem.py `orig` / `perturb` -> Solver.run (adFVM/solver.py:211-417): nSteps calls of `primal`,
                 state fed back, objective accumulated; result = sum / (nSteps - avgStart) (apps/problem.py:47-49)
  run_adjoint  = apps/adjoint.py Adjoint.run (:176-401): per checkpoint block a forward sweep that keeps every state,
                 then the block's steps backwards through `primal_grad` (adjoint fields fed back, obja = 1, sample every
                 step with return_static + zero_static), sensitivity = sum(paramGradient * perturbation) per step
                 (cmesh.computeSensitivity), result = sum / nSteps
The functions take any pair of callables with the reference's calling convention (adfvm_b200.function objects on the
device or the CPU simulator, the oracle, oracle/refgraph), so the same loop checks all of them against the anchors
of BASELINE.md section 5 recorded from the unmodified reference (tests/golden/anchor_*.{npz,json}).
Fixed time step, parameters = 'source', one perturbation, avgStart = 0 (the anchors' settings)."""
import contextlib

import numpy as np


@contextlib.contextmanager
def source_terms(case, source):
    """the case's source-term inputs for the duration of a run (None = zero, the drivers' default `source` lambda,
    apps/problem.py:16)"""
    C = case.mesh.nInternalCells
    src = [np.zeros((C, 1)), np.zeros((C, 3)), np.zeros((C, 1))] if source is None else source
    saved, case.source = case.source, [np.ascontiguousarray(s, case.dtype) for s in src]
    case._static = None
    try:
        yield
    finally:
        case.source = saved
        case._static = None


def run_primal(primal, case, nSteps, source=None, state=None, keep=False):
    """source: the three source-term arrays (None = zero); returns (result, final state, per-step objectives[, states])"""
    with source_terms(case, source):
        state = list(case.state if state is None else state)
        states, series = [state], []
        for i in range(nSteps):
            out = primal(*case.inputs(state), replace_reusable=(i == 0), return_reusable=True, replace_static=(i == 0))
            state = [np.array(o, copy=True) for o in out[:3]]
            series.append(float(out[4][0, 0]))
            if keep:
                states.append(state)
    res = sum(series) / nSteps
    return (res, state, series, states) if keep else (res, state, series)


def run_adjoint(primal, primal_grad, case, nSteps, writeInterval, perturbation, scaling=0.0):
    """returns (result, final adjoint fields, per-step sensitivities in the order the reference appends them).
    scaling: adjParams[0]; primal_grad is then the viscous function (viscousInterval = 1, apps/adjoint.py:250,288-289)"""
    C = case.mesh.nInternalCells
    adj = [np.zeros((C, 1), case.dtype), np.zeros((C, 3), case.dtype), np.zeros((C, 1), case.dtype)]
    # checkpoint states at multiples of writeInterval (the reference reads them back from disk)
    _, _, _, states = run_primal(primal, case, nSteps, keep=True)
    result, sens = 0.0, []
    for checkpoint in range(nSteps // writeInterval):
        primalIndex = nSteps - (checkpoint + 1) * writeInterval
        _, _, _, block = run_primal(primal, case, writeInterval, state=states[primalIndex], keep=True)
        with source_terms(case, None):          # the adjoint run linearises about the unperturbed trajectory
            for step in range(writeInterval):
                adjointIndex = writeInterval - 1 - step
                out = primal_grad(*case.adjoint_inputs(block[adjointIndex], adj, obja=1.0, dtca=0.0, scaling=scaling),
                                  return_static=True, zero_static=True, return_reusable=True, replace_reusable=False,
                                  replace_static=(step == 0))
                adj = [np.array(o, copy=True) for o in out[:3]]
                s = sum(float((np.asarray(g, np.float64) * np.asarray(p, np.float64)).sum()) for g, p in zip(out[3:6], perturbation))
                result += s
                sens.append(s)
    return result / nSteps, adj, sens
