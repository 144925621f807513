"""Test helper: minimal stand-ins for the objects of the reference's adpy trace (class names and attributes as in
adpy/adpy/scalar.py, tensor.py, variable.py), enough to hand-build the graph of a case-file objective where the reference's
front-end is not available (the GPU box). adfvm_b200.adpy_objective inspects traces by class name and attribute only."""


class _S:
    def __init__(self, *args):
        self.args = tuple(args)


class Scalar(_S): pass
class IntegerScalar(_S): pass
class ConstantOp(_S):
    def __init__(self, c): self.args = (); self.constant = c
class AddOp(_S): pass
class SubOp(_S): pass
class MulOp(_S): pass
class DivOp(_S): pass
class Extract(_S): pass
class Reduce(_S):
    def __init__(self, opType, x): self.opType = opType; self.args = (x,)


class Tensor:
    def __init__(self, n, integer=False, cell=False):
        self.scalars = [(IntegerScalar if integer else Scalar)() for _ in range(n)]
        self.cellTensor = cell


class Variable:
    _n = 0

    def __init__(self, shape, name=None, index=0, args=()):
        Variable._n += 1
        self.shape, self.index, self.args = shape, index, args
        self.name = name or "Variable_%d" % Variable._n

    def ref(self, index=0):
        return Variable(self.shape, self.name, index, (self,))


class Zeros(Variable): pass


class TensorFunction:
    def __init__(self, inputs, outputs):
        self._inputTensors, self._outputTensors = inputs, outputs


class _Out:
    def __init__(self, scalars): self.scalars = scalars


class TensorFunctionOp:
    def __init__(self, func, args, outputs, indices):
        self.func, self.indices = func, indices
        self.args = tuple(args) + tuple(outputs)
        self.outputs = tuple(Variable(o.shape, o.name, 0, (self,)) for o in outputs)


class Function:
    def __init__(self, inputs): self._inputs = inputs


def patch_pressure_force(n_inputs_vars, fields, areas_var, neighbour_var, startFace, nFaces):
    """objective = sum over the faces [startFace, startFace+nFaces) of p[neighbour[f]] * areas[f] (templates/forwardStep.py:8-28)"""
    tU, tT, tp = Tensor(3, cell=True), Tensor(1, cell=True), Tensor(1, cell=True)
    tA, tN = Tensor(1), Tensor(1, integer=True)
    val = MulOp(Extract(tp.scalars[0], tN.scalars[0]), tA.scalars[0])
    f = TensorFunction([tU, tT, tp, tA, tN], [_Out([Reduce("sum", val)])])
    obj = Zeros((1, 1))
    op = TensorFunctionOp(f, [fields[0], fields[1], fields[2], areas_var.ref(startFace), neighbour_var.ref(startFace)], [obj], nFaces)
    return op.outputs[0]


def cell_TV(fields, volumes_var, nInternalCells):
    """objective = sum over the internal cells of T * V (templates/box.py style, gen_golden OBJ_CELL_TV)"""
    tU, tT, tp, tV = Tensor(3), Tensor(1), Tensor(1), Tensor(1)
    f = TensorFunction([tU, tT, tp, tV], [_Out([Reduce("sum", MulOp(tT.scalars[0], tV.scalars[0]))])])
    obj = Zeros((1, 1))
    op = TensorFunctionOp(f, [fields[0], fields[1], fields[2], volumes_var], [obj], nInternalCells)
    return op.outputs[0]
