"""Drop-in overlay for the reference's `adpy` package: NO edit of the reference tree, only PYTHONPATH.

    PYTHONPATH=<this repo>/adfvm_b200/dropin:<reference>/adpy:<reference>  python <reference>/apps/problem.py <case> ...
    PYTHONPATH=...same...                                                   python <reference>/apps/adjoint.py <case> ...

This package shadows `adpy` and re-exports the reference's own modules unchanged (its directory is appended to this
package's `__path__`, so `adpy.tensor`, `adpy.variable`, `adpy.scalar`, `adpy.config` ARE the reference's files): the case
file, solver.py, density.py, BCs.py, the objective and `mu` lambdas are traced by the reference's own front-end exactly as
before. What changes is what happens at `Function.compile` (adpy/adpy/variable.py:545-597): instead of writing C++/CUDA and
calling the compiler, the step functions are bound to the hand-written sm_100a kernels of adfvm_b200:

  primal       -> adfvm_b200.function.PrimalFunction    (adFVM/density.py:101-105, adFVM/solver.py:312-323)
  primal_grad  -> adfvm_b200.function.AdjointFunction   (apps/adjoint.py:94-126, 268-291)
  init         -> adfvm_init_fields                      (adFVM/density.py:64-80: ghost-filled conservative fields for writing)

The case file's objective - arbitrary adpy-DSL code - is taken from the trace (adfvm_b200.adpy_objective) and evaluated /
differentiated on the device arrays; nothing has to be declared. `primal_grad_viscous` (a case file with adjParams =
[scaling, type, None]: adjoint artificial viscosity, apps/adjoint.py:127-141) -> AdjointFunction.viscous(type). Anything else a
case asks of the generated module (`compute_energy` on the device, dynamic meshes) raises NotImplementedError naming it.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_real = None
for _p in sys.path:
    _cand = os.path.join(_p or ".", "adpy")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and os.path.exists(os.path.join(_cand, "variable.py")):
        _real = os.path.abspath(_cand)
        break
if _real is None:
    raise ImportError("adfvm_b200 drop-in: the reference's adpy package was not found on sys.path after %s" % _here)
__path__.append(_real)

from . import tensor as _tensor          # noqa: E402  (the reference's modules, found through __path__)
from . import variable as _variable      # noqa: E402

_STATE = {"fields": [], "module": None}


def _solver():
    f = sys.modules.get("adFVM.field")
    return getattr(getattr(f, "Field", None), "solver", None) if f else None


# ---- 1. remember which Variables the case file's objective received: the frame of `solver.objective` is on the stack while its
# kernels are traced (adFVM/density.py:354-355)
_Kernel = _tensor.Kernel


def Kernel(func):
    param = _Kernel(func)

    def ParamFunc(indices=None, outputs=None):
        inner = param(indices, outputs)

        def Func(*args, **kwargs):
            s = _solver()
            obj = getattr(s, "objective", None)
            code = getattr(obj, "__code__", None)
            if code is not None:
                fr = sys._getframe(1)
                while fr is not None:
                    if fr.f_code is code:
                        fields = fr.f_locals.get(code.co_varnames[0])
                        if isinstance(fields, (list, tuple)) and len(fields) == 3 and \
                                not any(t[0] is fields[0] for t in _STATE["fields"]):
                            _STATE["fields"].append(list(fields))       # one triple per RK stage the objective is traced at
                        break
                    fr = fr.f_back
            return inner(*args, **kwargs)
        return Func
    return ParamFunc


_tensor.Kernel = Kernel


# ---- 2. the "generated module": what Function.__call__ looks its functions up in (adpy/adpy/variable.py:533-538)
class _Module:
    def __init__(self, solver):
        import numpy as np
        root = os.path.dirname(os.path.dirname(os.path.dirname(_here)))
        if root not in sys.path:
            sys.path.append(root)
        from adfvm_b200 import function as b200, adpy_objective
        from adFVM import config
        self.solver = solver
        objective = {"kind": "none"}
        if getattr(solver, "objective", None) is not None:
            if not _STATE["fields"]:
                raise NotImplementedError("the objective of the case file did not trace any kernel on the fields it was given")
            traced = adpy_objective.TracedObjective(solver.map, _STATE["fields"], solver.map._outputs[4])
            objective = {"kind": "traced", "traced": traced}
        spec = b200.spec_from_solver(solver, objective)
        main = sys.modules.get("__main__")
        par = getattr(main, "parameters", None)              # apps/adjoint.py:101-120 reads the case file's `parameters`
        if isinstance(par, (list, tuple)) and len(par) == 1:
            par = par[0]
        if par not in (None, [], ()):
            spec["parameters"] = par if isinstance(par, str) else tuple(par)
        device = int(os.environ.get("ADFVM_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        lib = None
        if os.environ.get("ADFVM_DROPIN_LIB"):               # tests: the CPU simulator of the device code
            from adfvm_b200 import _lib
            lib = _lib.Lib(os.environ["ADFVM_DROPIN_LIB"])
        self.primal_f = b200.PrimalFunction(spec, config.precision, device=device, lib=lib)
        self.grad_f = self.primal_f.grad()
        self.visc_f, self.viscous_calls = None, 0
        self.np = np

    def initialize(self, *args, **kwargs):
        return None

    def primal(self, *args, **kwargs):
        return self.primal_f(*args, **kwargs)

    def primal_grad(self, *args, **kwargs):
        return self.grad_f(*args, **kwargs)

    def primal_grad_viscous(self, *args, **kwargs):
        """`adjoint.viscousMap` (apps/adjoint.py:127-141,288-289): the viscosity type is adjParams[1] of the case file, which
        apps/problem.py keeps as a module global (apps/problem.py:22,26-33); the scaling arrives as the last positional input"""
        if self.visc_f is None:
            prob = sys.modules.get("problem") or sys.modules.get("__main__")
            adj = getattr(prob, "adjParams", None) or getattr(sys.modules.get("__main__"), "adjParams", None)
            if not adj or adj[1] is None:
                raise NotImplementedError("primal_grad_viscous called but the case file's adjParams names no viscosity type")
            self.visc_f = self.grad_f.viscous(adj[1])
        self.viscous_calls += 1
        return self.visc_f(*args, **kwargs)

    def init(self, *args, **kwargs):
        if not self.primal_f.c.static_loaded:
            # `init` may come before the first step (fields written at start-up): load the static inputs from the positional list
            # Solver.run would pass to `primal` (adFVM/solver.py:312-317)
            s, mesh, np = self.solver, self.solver.mesh, self.np
            from adFVM import config
            full = list(args[:3]) + [np.zeros((1, 1), config.precision)] + mesh.getTensor() + mesh.getScalar() + \
                [x[1] for x in s.sourceTerms] + s.getBoundaryTensor(1) + [x[1] for x in s.extraArgs]
            self.primal_f.set_state(*full)
        return self.primal_f.init_fields(*args)

    def __getattr__(self, name):
        def missing(*a, **k):
            raise NotImplementedError("function %r of the generated module is outside the residual / adjoint hot path served by "
                                      "adfvm_b200 (SURVEY section 8)" % name)
        return missing


Function = _variable.Function


def _compile(cls, case="./", init=True, replace=True, compiler_args={}):
    s = _solver()
    if s is None or not hasattr(s, "map"):
        raise NotImplementedError("adfvm_b200 drop-in: Function.compile outside an adFVM solver")
    cls._module = _Module(s)
    _STATE["module"] = cls._module
    if init:
        cls.initialize()
    cls._init = False


def _initialize(cls, *args, **kwargs):
    return None


Function.compile = classmethod(_compile)
Function.initialize = classmethod(_initialize)
Function.createCodeDir = classmethod(lambda cls, case, replace=True: None)
