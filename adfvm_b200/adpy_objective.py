"""Any case-file objective, evaluated and differentiated on the device arrays.

The reference takes the objective as adpy-DSL code in the case file (`objective(fields, solver)`, e.g.
templates/cylinder_test.py:9-36, adFVM/objectives/vane.py:83-140): `tensor.Kernel(f)(n, outputs)(*args)` traces the Python
function `f` with the reference's own front-end (adpy/adpy/tensor.py:444-485) into a DAG of scalar ops
(adpy/adpy/scalar.py:119-320) inside the `primal` function graph, and `Function.grad` differentiates it together with the
residual. Here the residual is hand-written CUDA; the objective sub-graph - whatever the case file wrote - is taken from that
same trace and interpreted with torch tensor ops on the stage-1 primitives living in HBM, its reverse mode by torch autograd:

  * `TracedObjective(function, field_vars, obj_var)` collects the ops between the objective's inputs (the U, T, p Variables
    handed to `objective()` and inputs of the `primal` function: mesh arrays, BC arrays, extraArgs) and its output;
  * `bind(inputs)` attaches the positional arguments of a `primal` call (static arrays are uploaded once);
  * `__call__(Q, want_seed, obja, Qseed)` is what the native library calls on the stage-1 state (include/adfvm_b200.h
    adfvm_set_objective_callback): value = the rank-local objective, Qseed = obja * dJ/dQ.

The scalar-op semantics follow the reference: `abs` -> sign with abs'(x<0) = -1 else +1 (scalar.py:209-212, also torch's except
at exactly 0), `switch` passes no gradient to its condition (:245-269), max/min reductions carry no gradient (:300-311),
`mpi_allreduce` sums over the ranks forward and copies the adjoint backward (adFVM/cpp/parallel.cpp:214-239). An allreduce that
ends the objective is left to the library, which sums the objective over the ranks like for its native objectives.
Nothing of adpy is imported here: the traced objects are inspected by attribute and class name.
"""
from __future__ import annotations

import numpy as np


def _cls(o):
    return type(o).__name__


class _AllReduce:
    """sum over ranks; backward = identity (the reference's mpi_allreduce_grad)"""
    _fn = None

    @classmethod
    def apply(cls, x):
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return x
        if cls._fn is None:
            class F(torch.autograd.Function):
                @staticmethod
                def forward(ctx, t):
                    t = t.clone()
                    dist.all_reduce(t)
                    return t

                @staticmethod
                def backward(ctx, g):
                    return g
            cls._fn = F
        return cls._fn.apply(x)


class TracedObjective:
    def __init__(self, function, field_vars, obj_var):
        """function: the reference's `primal` Function (inputs in call order); field_vars: the [U, T, p] Variables the case
        file's objective received - or a list of such triples, one per RK stage the objective was traced at (the step function
        keeps the stage-1 one, adFVM/density.py:412-413): the triple the objective's graph reaches is used; obj_var: the Variable
        it returned"""
        self.in_names = [v.name for v in function._inputs]
        self.n_inputs = len(self.in_names)
        self.in_kinds = [_cls(v) for v in function._inputs]
        candidates = field_vars if isinstance(field_vars[0], (list, tuple)) else [field_vars]
        candidates = [[v.name for v in t] for t in candidates]
        self.final_allreduce = False
        # ---- ops between the objective's inputs and its output, in evaluation order
        known = set(self.in_names) | set(n for t in candidates for n in t)
        order, seen, hit = [], set(), set()

        def producer(var):                              # walk reference chains (x[offset]) down to the producing op
            while var.args and _cls(var.args[0]) not in ("TensorFunctionOp", "ExternalFunctionOp"):
                var = var.args[0]
            return var.args[0] if var.args else None

        def visit(var):
            stack = [(var, False)]
            while stack:
                v, done = stack.pop()
                if v.name in known and not done:
                    hit.add(v.name)
                    continue
                op = producer(v)
                if op is None:
                    if _cls(v) == "Zeros" or v.name in known:
                        continue
                    raise NotImplementedError("objective depends on %s, which is neither a field handed to objective() nor an "
                                              "input of the step function" % v.name)
                if done:
                    if id(op) not in seen:
                        seen.add(id(op)); order.append(op)
                    continue
                if id(op) in seen:
                    continue
                stack.append((v, True))
                for a in op.args:
                    stack.append((a, False))
        visit(obj_var)
        used = [t for t in candidates if hit & set(t)]
        if len(used) > 1:
            raise NotImplementedError("objective mixes the fields of several RK stages")
        self.field_names = used[0] if used else candidates[0]
        self.ops = order
        self.out_name = obj_var.name
        if order and _cls(order[-1]) == "ExternalFunctionOp" and order[-1].name == "Function_mpi_allreduce":
            self.final_allreduce = True
            self.ops = order[:-1]
            self.out_name = order[-1].args[0].name
        for op in self.ops:
            if _cls(op) == "ExternalFunctionOp" and op.name != "Function_mpi_allreduce":
                raise NotImplementedError("external op %s inside an objective" % op.name)
        self.bound = None
        self.device = None
        self.error = None

    # ---- runtime
    def bind(self, inputs, device, dtype, row_of_cell, mesh_param=False):
        """inputs: positional list of a `primal` call; row_of_cell: long tensor, device row of every reference cell index
        (internal cells permuted into tile order, ghost rows in place); mesh_param: the adjoint's parameter block is the ten
        mesh metric arrays (parameters = 'mesh', apps/adjoint.py:105-107) - the objective's own dependence on them is then
        differentiated too and accumulated in `mesh_grad` (reference numbering, like the arrays themselves)"""
        import torch
        self.device, self.dtype = device, dtype
        self.mesh_names = self.in_names[4:14] if mesh_param else []
        self.mesh_grad = None
        vals = {}
        for name, kind, a in zip(self.in_names, self.in_kinds, inputs):
            if isinstance(a, np.ndarray):
                t = torch.as_tensor(a, device=device)
                vals[name] = t.to(dtype) if t.dtype.is_floating_point else t.long()
                if vals[name].dim() == 1:
                    vals[name] = vals[name].reshape(-1, 1)
            else:
                vals[name] = int(a)
        for name in self.mesh_names:
            vals[name].requires_grad_(True)
        self.bound = vals
        self.rows = row_of_cell

    def _offset(self, idx, vals):
        return idx if isinstance(idx, (int, np.integer)) else int(vals[idx.name])

    def _dim(self, d, vals):
        return d if isinstance(d, (int, np.integer)) else int(vals[d.name])

    def evaluate(self, Qt):
        """Qt: torch tensor [5][stride] (device layout); returns the objective as a 0-d tensor (graph attached if Qt requires grad)"""
        import torch
        vals = dict(self.bound)
        U = Qt[0:3][:, self.rows].t()
        vals[self.field_names[0]] = U
        vals[self.field_names[1]] = Qt[3][self.rows].reshape(-1, 1)
        vals[self.field_names[2]] = Qt[4][self.rows].reshape(-1, 1)
        for op in self.ops:
            if _cls(op) == "ExternalFunctionOp":
                nin = len(op.args) - len(op.outputs)
                for a, o in zip(op.args[:nin], op.outputs):
                    vals[o.name] = _AllReduce.apply(vals[a.name])
                continue
            self._kernel(op, vals, torch)
        return vals[self.out_name].reshape(-1)[0]

    def _kernel(self, op, vals, torch):
        f = op.func
        nin = len(f._inputTensors)
        n = self._dim(op.indices, vals)
        actual_in, actual_out = op.args[:nin], op.args[nin:]
        base = {}                                       # formal scalar -> (buffer [rows, comps], offset, component)
        for formal, act in zip(f._inputTensors, actual_in):
            buf = vals[act.name]
            buf2 = buf.reshape(buf.shape[0], -1)
            off = self._offset(act.index, vals)
            for j, s in enumerate(formal.scalars):
                base[s] = (buf2, off, j, formal.cellTensor)
        ar = None
        memo = {}

        def val(s):
            stack = [s]
            while stack:
                x = stack[-1]
                if x in memo:
                    stack.pop(); continue
                c = _cls(x)
                if x in base:
                    buf, off, j, cellT = base[x]
                    if cellT:
                        memo[x] = None                  # only reachable through Extract / Singular
                    else:
                        memo[x] = buf[off:off + n, j]
                    stack.pop(); continue
                if c == "ConstantOp":
                    memo[x] = x.constant; stack.pop(); continue
                if c == "IndexOp":
                    memo[x] = torch.arange(n, device=self.device); stack.pop(); continue
                if c in ("Extract", "Singular"):
                    if x.args[0] not in base:
                        raise NotImplementedError("%s of a computed value in an objective kernel" % c)
                    need = [a for a in x.args[1:] if a not in memo]
                else:
                    need = [a for a in x.args if a not in memo]
                if need:
                    stack.extend(need); continue
                a = [memo.get(q) for q in x.args]
                if c == "AddOp": r = a[0] + a[1]
                elif c == "SubOp": r = a[0] - a[1]
                elif c == "MulOp": r = a[0] * a[1]
                elif c == "DivOp": r = a[0] / a[1]
                elif c == "PowerOp": r = a[0] ** a[1]
                elif c == "LessThanOp": r = a[0] < a[1]
                elif c == "NegOp": r = -a[0]
                elif c == "AbsOp": r = abs(a[0]) if not torch.is_tensor(a[0]) else torch.abs(a[0])
                elif c == "SqrtOp": r = a[0] ** 0.5 if not torch.is_tensor(a[0]) else torch.sqrt(a[0])
                elif c == "InvertOp": r = ~a[0] if torch.is_tensor(a[0]) else (not a[0])
                elif c == "ConditionalOp":
                    cond = a[0] if torch.is_tensor(a[0]) else torch.full((n,), bool(a[0]), device=self.device)
                    t1 = a[1] if torch.is_tensor(a[1]) else torch.full((n,), float(a[1]), device=self.device, dtype=self.dtype)
                    t2 = a[2] if torch.is_tensor(a[2]) else torch.full((n,), float(a[2]), device=self.device, dtype=self.dtype)
                    r = torch.where(cond, t1, t2)
                elif c == "Extract":
                    buf, off, j, _ = base[x.args[0]]
                    r = buf[off + a[1], j]
                elif c == "Singular":
                    buf, off, j, _ = base[x.args[0]]
                    r = buf[off, j]
                else:
                    raise NotImplementedError("scalar op %s in an objective kernel" % c)
                memo[x] = r
                stack.pop()
            return memo[s]

        def vec(v):
            return v if torch.is_tensor(v) else torch.full((n,), float(v), device=self.device, dtype=self.dtype)

        for formal, act, ref in zip(f._outputTensors, actual_out, op.outputs):
            shape0 = self._dim(act.shape[0], vals)
            width = int(np.prod(act.shape[1:]))
            seed = vals.get(act.name)               # kernels ACCUMULATE into the output they are given (Zeros unless produced earlier)
            out = torch.zeros((shape0, width), device=self.device, dtype=self.dtype) if seed is None else seed.reshape(shape0, width)
            off = self._offset(act.index, vals)
            delta = torch.zeros((shape0, width), device=self.device, dtype=self.dtype)
            for j, s in enumerate(formal.scalars):
                if s is None:
                    continue
                c = _cls(s)
                if c == "Reduce":
                    v = vec(val(s.args[0]))
                    if s.opType == "sum":
                        delta[off, j] = v.sum()
                    else:                               # max / min: no gradient in the reference (scalar.py:306-311)
                        delta[off, j] = (v.max() if s.opType == "max" else v.min()).detach()
                elif c == "Collate":
                    col = torch.zeros(shape0, device=self.device, dtype=self.dtype)
                    for k in range(len(s.args) // 2):
                        col = col.index_add(0, off + val(s.args[2 * k + 1]), vec(val(s.args[2 * k])))
                    delta[:, j] = col
                else:
                    delta[off:off + n, j] = vec(val(s))
            vals[ref.name] = (out + delta).reshape((shape0,) + tuple(int(d) for d in act.shape[1:]))

    # ---- the native callback (include/adfvm_b200.h adfvm_objective_fn)
    def __call__(self, Q, want_seed, obja, Qseed):
        """Q / Qseed: torch tensors [5][stride] over the library's device memory"""
        import torch
        if not want_seed:
            with torch.no_grad():
                return float(self.evaluate(Q).detach())
        Qt = Q.detach().clone().requires_grad_(True)
        J = self.evaluate(Qt)
        mesh = [self.bound[n] for n in self.mesh_names]
        grads = torch.autograd.grad(J, [Qt] + mesh, allow_unused=True)
        if grads[0] is not None:
            Qseed.add_(grads[0], alpha=obja)
        if mesh:
            if self.mesh_grad is None:
                self.mesh_grad = [torch.zeros_like(t) for t in mesh]
            for acc, g in zip(self.mesh_grad, grads[1:]):
                if g is not None:
                    acc.add_(g, alpha=obja)
        return float(J.detach())

    def take_mesh_grad(self, zero):
        """accumulated obja * dJ/d(mesh arrays) as numpy arrays in Mesh.gradFields order (None if nothing was accumulated)"""
        if self.mesh_grad is None:
            return None
        out = [g.detach().cpu().numpy().copy() for g in self.mesh_grad]
        if zero:
            for g in self.mesh_grad:
                g.zero_()
        return out
