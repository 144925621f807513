"""Host-side mirror of the reference's compiled-function objects for the residual hot path.

In the reference `solver.map` / `adjoint.map` are `adpy.variable.Function` instances whose `__call__`
forwards a long positional list of numpy arrays / ints plus option kwargs to the generated
`graph_N.Function_primal` / `Function_primal_grad` (adpy/adpy/variable.py:533-538). `PrimalFunction` and
`AdjointFunction` below are callable with exactly that positional list and those kwargs and return exactly
those tuples, but land in the C ABI of include/adfvm_b200.h (hand-written sm_100a kernels).

What the reference bakes into generated code at compile time (patch types, BC classes, gas constants, the
objective) is passed once as a `spec` dict — `spec_from_solver(primal)` reads it off a reference `RCF`
object, tests read it from the golden fixtures:

  spec = {Cp, gamma, Pr, mu: {law: sutherland|constant, value}, riemannSolver, boundaryRiemannSolver,
          timeIntegrator: 'SSPRK', sortedPatches: [...],
          patches: [{name, type, startFace, nFaces, cellStartFace, neighbourPatch?, myProcNo?, neighbProcNo?, tag?}],
          BCs: {U|T|p: {patch: {type, keys}}}, objective: {kind: none|cell_TV|patch_pA|drag, patch?, direction?}}

Positional layout (adFVM/solver.py:312-317, adFVM/mesh.py:870-881, adFVM/field.py:141-146,
adFVM/density.py:108-110): 0-2 state | 3 dt | 4-13 gradFields | 14-18 intFields | 19-26 constants |
3 ints per sorted local patch | 3 source arrays | BC value arrays (U,T,p then gradient fields) | extraArgs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L

GRAD_FIELDS = ["areas", "volumesL", "volumesR", "weights", "deltas", "normals", "deltasUnit",
               "linearWeights", "quadraticWeights", "volumes"]
GRAD_FIELD_DIMS = [(1,), (1,), (1,), (1,), (1,), (3,), (3,), (2,), (2, 3), (1,)]
INT_FIELDS = ["owner", "neighbour", "cellFaces", "cellNeighbours", "cellOwner"]
CONSTANTS = ["nCells", "nFaces", "nInternalCells", "nInternalFaces",
             "nLocalCells", "nRemoteCells", "nLocalFaces", "nGhostCells"]

_MESH_TYPE = {"cyclic": L.PATCH_CYCLIC, "slidingPeriodic1D": None, "symmetryPlane": L.PATCH_SYMMETRY,
              "empty": L.PATCH_EMPTY, "characteristic": L.PATCH_CHARACTERISTIC, "processor": L.PATCH_PROCESSOR,
              "processorCyclic": L.PATCH_PROCESSOR_CYCLIC}
_BC_TYPE = {"cyclic": L.BC_CYCLIC, "zeroGradient": L.BC_ZEROGRADIENT, "empty": L.BC_ZEROGRADIENT,
            "inletOutlet": L.BC_ZEROGRADIENT, "fixedValue": L.BC_FIXEDVALUE, "symmetryPlane": L.BC_SYMMETRY,
            "slip": L.BC_SYMMETRY, "calculated": L.BC_CALCULATED, "CBC_UPT": L.BC_CBC_UPT,
            "CBC_TOTAL_PT": L.BC_CBC_TOTAL_PT, "processor": L.BC_PROCESSOR, "processorCyclic": L.BC_PROCESSOR}
_BC_KEY = {("U", "value"): (L.KEY_VALUE_U, 3), ("T", "value"): (L.KEY_VALUE_T, 1), ("p", "value"): (L.KEY_VALUE_P, 1),
           ("p", "U0"): (L.KEY_U0, 3), ("p", "T0"): (L.KEY_T0, 1), ("p", "p0"): (L.KEY_P0, 1),
           ("p", "Tt"): (L.KEY_TT, 1), ("p", "pt"): (L.KEY_PT, 1), ("p", "direction"): (L.KEY_DIRECTION, 3)}
_OBJ_KIND = {"none": L.OBJ_NONE, "cell_TV": L.OBJ_CELL_TV, "patch_pA": L.OBJ_PATCH_PA, "drag": L.OBJ_DRAG, "cell_T": L.OBJ_CELL_T}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """One page-locked host block; numpy arrays made from it keep it alive (array interface base), and the block
    goes back to its pool when the last of them is released."""

    def __init__(self, pool, ptr, nbytes):
        self.pool, self.ptr, self.nbytes = pool, ptr, nbytes
        self.__array_interface__ = {"data": (ptr, False), "shape": (nbytes,), "typestr": "|u1", "version": 3}

    def __del__(self):
        pool = self.pool
        if pool is not None:
            pool._release(self.ptr, self.nbytes)


class _PinnedPool:
    """Result arrays of a call (reference: fresh `new T[]` per output, adpy/adpy/cpp/include/interface.hpp:54-80)
    are fresh numpy arrays over recycled page-locked blocks: the device-to-host copies run asynchronously at the
    full PCIe rate instead of staging through pageable memory."""

    def __init__(self, lib, keep_bytes=None):
        # recycled blocks are kept up to a byte budget (a checkpoint block of the reference's adjoint driver holds one state per
        # step alive, apps/adjoint.py:217; allocating page-locked memory is slow, ~0.5 s per GB)
        if keep_bytes is None:
            keep_bytes = int(float(os.environ.get("ADFVM_PINNED_POOL_GB", "16")) * 2 ** 30)
        self.lib, self.free, self.keep_bytes, self.pooled = lib, {}, keep_bytes, 0

    def empty(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        nbytes = max(64, (n + 63) // 64 * 64)
        lst = self.free.get(nbytes)
        if lst:
            ptr = lst.pop()
            self.pooled -= nbytes
        else:
            p = C.c_void_p()
            self.lib.check(self.lib.dll.adfvm_host_alloc(C.byref(p), nbytes))
            ptr = p.value
        block = _PinnedBlock(self, ptr, nbytes)
        return np.asarray(block)[:n].view(dtype).reshape(shape)

    def _release(self, ptr, nbytes):
        lst = self.free.setdefault(nbytes, [])
        if self.pooled + nbytes <= self.keep_bytes or nbytes <= 4096:
            lst.append(ptr)
            self.pooled += nbytes
        else:
            try:
                self.lib.dll.adfvm_host_free(C.c_void_p(ptr))
            except Exception:
                pass

    def close(self):
        for lst in self.free.values():
            for ptr in lst:
                try:
                    self.lib.dll.adfvm_host_free(C.c_void_p(ptr))
                except Exception:
                    pass
        self.free, self.pooled = {}, 0


class _Context:
    """Owns one adfvm_ctx: device-resident mesh, BC, source and state of one rank."""

    def __init__(self, spec, precision, device, stream, lib, tile_cells=None):
        self.spec = spec
        self.lib = lib or L.default_lib()
        self.dtype = np.dtype(precision)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("precision must be float32 or float64")
        if spec.get("timeIntegrator", "SSPRK") != "SSPRK":
            raise NotImplementedError("only the 3-stage SSPRK integrator is supported (the reference's euler "
                                      "integrator cannot compile either, SURVEY §3.1)")
        self.ctx = C.c_void_p()
        self.device, self.stream = int(device), stream
        self.lib.check(self.lib.dll.adfvm_create(C.byref(self.ctx), int(device), self.dtype.itemsize,
                                                  C.c_void_p(stream) if stream else None))
        mu = spec["mu"]
        law = L.MU_SUTHERLAND if mu["law"] == "sutherland" else L.MU_CONSTANT
        self.lib.check(self.lib.dll.adfvm_set_physics(
            self.ctx, float(spec["gamma"]), float(spec["Cp"]), float(spec["Pr"]), law, float(mu.get("value", 0.)),
            L.RIEMANN[spec["riemannSolver"]], L.RIEMANN[spec["boundaryRiemannSolver"]]))
        if tile_cells:
            self.lib.check(self.lib.dll.adfvm_set_tile_cells(self.ctx, int(tile_cells)))
        self.sorted = list(spec["sortedPatches"])
        byname = {p["name"]: p for p in spec["patches"]}
        remote = [p["name"] for p in spec["patches"] if p["type"] in ("processor", "processorCyclic")]
        self.patch_names = self.sorted + [n for n in remote if n not in self.sorted]
        self.patch = byname
        self.static_loaded = False
        self.sizes = None
        self.pool = _PinnedPool(self.lib)

    def close(self):
        if self.ctx:
            self.pool.close()
            self.lib.dll.adfvm_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- argument checking in the spirit of getArray (adpy/adpy/cpp/include/interface.hpp:29-52)
    def _arr(self, a, shape_tail, what, rows=None):
        if not isinstance(a, np.ndarray):
            raise TypeError("%s: expected a numpy array" % what)
        if a.dtype != self.dtype:
            raise TypeError("%s: dtype %s, expected %s" % (what, a.dtype, self.dtype))
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("%s: array must be C-contiguous" % what)
        if tuple(a.shape[1:]) != tuple(shape_tail):
            raise ValueError("%s: trailing dims %s, expected %s" % (what, a.shape[1:], tuple(shape_tail)))
        if rows is not None and a.shape[0] != rows:
            raise ValueError("%s: %d rows, expected %d" % (what, a.shape[0], rows))
        return a

    def _iarr(self, a, what, rows, cols=None):
        if not isinstance(a, np.ndarray) or a.dtype != np.int32 or not a.flags["C_CONTIGUOUS"]:
            raise TypeError("%s: expected a C-contiguous int32 array" % what)
        if a.shape[0] != rows or (cols is not None and (a.ndim != 2 or a.shape[1] != cols)):
            raise ValueError("%s: bad shape %s" % (what, a.shape))
        return a

    def parse(self, inputs):
        """Split the positional list; returns dict with state, dt, mesh arrays, sizes, source, bc arrays, rest."""
        k = 4
        if len(inputs) < 27:
            raise TypeError("expected at least 27 positional inputs, got %d" % len(inputs))
        out = {"state": inputs[0:3], "dt": inputs[3]}
        mesh = {}
        for name in GRAD_FIELDS + INT_FIELDS:
            mesh[name] = inputs[k]; k += 1
        sizes = [int(inputs[k + i]) for i in range(8)]; k += 8
        triples = {}
        for pid in self.sorted:
            triples[pid] = (int(inputs[k]), int(inputs[k + 1]), int(inputs[k + 2])); k += 3
        source = inputs[k:k + 3]; k += 3
        bcs = []
        for field in ("U", "T", "p"):
            for pid in self.sorted:
                for key in self.spec["BCs"][field][pid]["keys"]:
                    bcs.append((field, pid, key, inputs[k])); k += 1
        # extraArgs of the case file (adFVM/solver.py:58,317): the objective's own inputs; the adjoint seeds follow them
        nx = int((self.spec.get("objective") or {}).get("nExtra", 0))
        extra = list(inputs[k:k + nx]); k += nx
        if (self.spec.get("objective") or {}).get("kind") == "traced":      # extraArgs: everything up to the adjoint seeds
            nx = int(self.spec["objective"]["traced"].n_inputs) - k
            extra = list(inputs[k:k + nx]); k += nx
        out.update(mesh=mesh, sizes=sizes, triples=triples, source=source, bcs=bcs, extra=extra, rest=list(inputs[k:]), inputs=list(inputs[:k]))
        return out

    def load_static(self, P):
        d = self.lib.dll
        sizes = P["sizes"]
        nCells, nFaces, C_, Fi = sizes[0], sizes[1], sizes[2], sizes[3]
        m = P["mesh"]
        rows = {"areas": nFaces, "volumesL": nFaces, "volumesR": Fi, "weights": nFaces, "deltas": nFaces,
                "normals": nFaces, "deltasUnit": nFaces, "linearWeights": nFaces, "quadraticWeights": nFaces,
                "volumes": C_}
        for name, dims in zip(GRAD_FIELDS, GRAD_FIELD_DIMS):
            self._arr(m[name], dims, "mesh." + name, rows[name])
        self._iarr(m["owner"], "mesh.owner", nFaces)
        self._iarr(m["neighbour"], "mesh.neighbour", nFaces)
        for name in ("cellFaces", "cellNeighbours", "cellOwner"):
            self._iarr(m[name], "mesh." + name, C_, 6)
        table = (L.Patch * len(self.patch_names))()
        index = {n: i for i, n in enumerate(self.patch_names)}
        for i, pid in enumerate(self.patch_names):
            p = self.patch[pid]
            if pid in P["triples"] and P["triples"][pid] != (p["startFace"], p["nFaces"], p["cellStartFace"]):
                raise ValueError("patch %s: (startFace,nFaces,cellStartFace) %s differs from the compiled spec"
                                 % (pid, P["triples"][pid]))
            if p["type"] == "slidingPeriodic1D":
                raise NotImplementedError("slidingPeriodic1D patches (dynamic mesh) are outside the hot path")
            t = table[i]
            t.startFace, t.nFaces, t.cellStartFace = p["startFace"], p["nFaces"], p["cellStartFace"]
            t.mesh_type = _MESH_TYPE.get(p["type"], L.PATCH_WALL)
            remote = p["type"] in ("processor", "processorCyclic")
            for fld, attr in (("U", "bc_U"), ("T", "bc_T"), ("p", "bc_p")):
                bct = p["type"] if remote else self.spec["BCs"][fld][pid]["type"]
                if bct not in _BC_TYPE:
                    raise NotImplementedError("boundary condition %r (patch %s) has no class in the reference "
                                              "BCs.py either" % (bct, pid))
                setattr(t, attr, _BC_TYPE[bct])
            t.neighbour_patch = index[p["neighbourPatch"]] if p["type"] == "cyclic" else -1
            t.peer_rank = int(p.get("neighbProcNo", -1)) if remote else -1
            t.tag = int(p.get("tag", 0))
        if (self.spec.get("parameters") or "source") == "mesh":      # must precede the mesh upload (see adfvm_b200.h)
            self.lib.check(d.adfvm_set_parameter_mesh(self.ctx))
        csizes = (C.c_int32 * 8)(*sizes)
        self.lib.check(d.adfvm_set_mesh(
            self.ctx, csizes, *[_ptr(m[n]) for n in GRAD_FIELDS], *[_ptr(m[n]) for n in INT_FIELDS],
            len(self.patch_names), table))
        self.sizes = sizes
        self.patch_index = index
        o = self.spec.get("objective") or {"kind": "none"}
        if o["kind"] == "plane_ptloss":
            # adFVM/objectives/vane.py: extraArgs = (nPlaneCells, cells[n,1], areas[n,1], weights of the `pressure` and
            # `suction` patches - those feed the heat-transfer term, whose coefficient b is 0 in the reference, :122)
            ex = P["extra"]
            if len(ex) < 3:
                raise TypeError("the cut-plane objective expects its extraArgs (nPlaneCells, cells, areas, ...)")
            n = int(ex[0])
            cells = self._iarr(np.ascontiguousarray(ex[1]), "objective cells", n, 1)
            areas = self._arr(ex[2], (1,), "objective areas", n)
            nrm = (C.c_double * 3)(*[float(x) for x in o.get("normal", (1., 0., 0.))])
            self.lib.check(d.adfvm_set_objective_plane(self.ctx, n, _ptr(cells), _ptr(areas), float(o.get("ptin", 175158.)),
                                                       nrm, float(o.get("scale", 0.4))))
        elif o["kind"] == "traced":
            self._install_traced_objective(o["traced"], P)
        else:
            if o["kind"] not in _OBJ_KIND:
                raise NotImplementedError("objective %r (supported: %s, plane_ptloss, traced)" % (o["kind"], ", ".join(_OBJ_KIND)))
            self.lib.check(d.adfvm_set_objective(self.ctx, _OBJ_KIND[o["kind"]],
                                                 index.get(o.get("patch"), 0), int(o.get("direction", 0))))
        self.load_replaceable(P)
        # parameter block of the adjoint: 'source' (default) or ('BCs', field, patch, key) (apps/adjoint.py:101-120)
        par = self.spec.get("parameters") or "source"
        self.param_bc = None
        self.param_mesh = par == "mesh"
        if par != "source" and par != "mesh":
            if not (isinstance(par, (list, tuple)) and len(par) == 4 and par[0] == "BCs"):
                raise NotImplementedError("parameters=%r: supported are 'source', 'mesh' and ('BCs', field, patch, key)" % (par,))
            _, field, pid, key = par
            if (field, key) not in _BC_KEY or key == "direction":
                raise NotImplementedError("no gradient with respect to BC input %s.%s" % (field, key))
            kid, dim = _BC_KEY[(field, key)]
            self.lib.check(d.adfvm_set_parameter_bc(self.ctx, self.patch_index[pid], kid))
            self.param_bc = (self.patch[pid]["nFaces"], dim)
        self.static_loaded = True

    # ---- an arbitrary case-file objective, traced by the reference's own front-end (adpy_objective.TracedObjective)
    def _install_traced_objective(self, traced, P):
        import torch
        d = self.lib.dll
        C_, nCells = self.sizes[2], self.sizes[0]
        perm = np.zeros(C_, np.int32)
        self.lib.check(d.adfvm_get_cell_perm(self.ctx, perm.ctypes.data_as(C.POINTER(C.c_int32))))
        rows = np.arange(nCells, dtype=np.int64)
        rows[perm] = np.arange(C_)                       # reference cell -> device row; ghost rows keep their place
        cuda = self.lib.is_cuda
        dev = torch.device("cuda", self.device) if cuda else torch.device("cpu")
        tdtype = torch.float64 if self.dtype == np.float64 else torch.float32
        traced.bind(P["inputs"], dev, tdtype, torch.as_tensor(rows, device=dev),
                    mesh_param=(self.spec.get("parameters") or "source") == "mesh")
        self.traced, self.traced_error = traced, None
        itemsize = self.dtype.itemsize

        def view(ptr, stride):
            if cuda:
                class _Dev:
                    __cuda_array_interface__ = {"shape": (5, int(stride)), "typestr": "<f%d" % itemsize, "data": (int(ptr), False), "version": 2}
                return torch.as_tensor(_Dev(), device=dev)
            buf = (C.c_char * (5 * int(stride) * itemsize)).from_address(int(ptr))
            return torch.from_numpy(np.frombuffer(buf, self.dtype).reshape(5, int(stride)))

        def callback(user, Q, stride, want_seed, obja, Qseed):
            try:
                Qt = view(Q, stride)
                Qs = view(Qseed, stride) if want_seed else None
                if cuda:
                    strm = torch.cuda.ExternalStream(self.stream, device=dev) if self.stream else torch.cuda.default_stream(dev)
                    with torch.cuda.stream(strm):
                        return traced(Qt, bool(want_seed), float(obja), Qs)
                return traced(Qt, bool(want_seed), float(obja), Qs)
            except BaseException as e:      # noqa: BLE001  (must not propagate through the C frames)
                self.traced_error = e
                return float("nan")
        self._objective_cb = L.OBJECTIVE_FN(callback)          # keep the trampoline alive as long as the context
        self.lib.check(d.adfvm_set_objective_callback(self.ctx, self._objective_cb, None))

    def check_traced(self):
        """re-raise what the objective callback caught during the last native call"""
        e, self.traced_error = getattr(self, "traced_error", None), None
        if e is not None:
            raise e

    def load_replaceable(self, P):
        """source terms + BC value arrays: static in the reference, but re-settable here (fixes the
        silent GPU-perturb bug noted at apps/problem.py:109-110)."""
        d = self.lib.dll
        C_ = P["sizes"][2]
        s = P["source"]
        self._arr(s[0], (1,), "source rho", C_); self._arr(s[1], (3,), "source rhoU", C_)
        self._arr(s[2], (1,), "source rhoE", C_)
        self.lib.check(d.adfvm_set_source(self.ctx, _ptr(s[0]), _ptr(s[1]), _ptr(s[2])))
        for field, pid, key, arr in P["bcs"]:
            if (field, key) not in _BC_KEY:
                raise NotImplementedError("BC input %s.%s" % (field, key))
            kid, dim = _BC_KEY[(field, key)]
            self._arr(arr, (dim,), "BC %s.%s.%s" % (field, pid, key), self.patch[pid]["nFaces"])
            self.lib.check(d.adfvm_set_bc_value(self.ctx, self.patch_index[pid], kid, _ptr(arr)))

    # ---- opt-in state cache (PrimalFunction(state_cache=n)): which returned array objects have a device copy
    def cache_register(self, arrs):
        """arrs: the three state arrays just returned (the resident state is theirs): make them read-only and file the copy"""
        import weakref
        if not self.state_cache:
            return
        if self.cache_next == 1:
            self.lib.check(self.lib.dll.adfvm_state_cache_reserve(self.ctx, self.state_cache))
        key, self.cache_next = self.cache_next, self.cache_next + 1
        self.lib.check(self.lib.dll.adfvm_state_cache_put(self.ctx, key))
        for a in arrs:
            a.flags.writeable = False
        ident = id(arrs[0])
        keys = self.cache_keys

        def gone(_ref, ident=ident, key=key):
            if keys.get(ident, (None,))[0] == key:
                keys.pop(ident, None)
        self.cache_keys[ident] = (key, [weakref.ref(arrs[0], gone), weakref.ref(arrs[1]), weakref.ref(arrs[2])])

    def cache_select(self, state):
        """True if `state` (rho, rhoU, rhoE) are exactly the array objects of a cached copy, untouched; the copy is then selected"""
        if not self.state_cache:
            return False
        ent = self.cache_keys.get(id(state[0]))
        if ent is None or any(r() is not a for r, a in zip(ent[1], state)) or any(a.flags.writeable for a in state):
            return False
        found = C.c_int32(0)
        self.lib.check(self.lib.dll.adfvm_state_cache_select(self.ctx, ent[0], C.byref(found)))
        return bool(found.value)

    def attach_comm(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        self.lib.check(self.lib.dll.adfvm_comm_init(self.ctx, buf, rank, nranks))


class PrimalFunction:
    """Drop-in for `solver.map` (Function('primal', ...), adFVM/density.py:101-105)."""
    name = "primal"
    defaultOptions = {"return_static": True, "zero_static": False, "replace_static": False,
                      "return_reusable": True, "replace_reusable": False}   # adpy/adpy/variable.py:282-287

    def __init__(self, spec, precision=np.float64, device=0, stream=None, lib=None, tile_cells=None, state_cache=0):
        """state_cache = n > 0 (opt-in, NOT the reference's semantics): the state arrays this function returns are READ-ONLY numpy
        arrays, and the library keeps device copies of the last n of them. When the very same array objects come back as the
        state of a `primal` (replace_reusable) or `primal_grad` call - what `Solver.run(mode='forward')` + `Adjoint.run` do with
        their `solutions` list (adFVM/solver.py:376-382, apps/adjoint.py:250-280) - and are still read-only, the device copy is
        used and the 5 C scalars do not travel up again. Anything else (other arrays, arrays made writeable again) is uploaded."""
        self.c = _Context(spec, precision, device, stream, lib, tile_cells)
        self.c.state_cache = int(state_cache)
        self.c.cache_keys, self.c.cache_next = {}, 1

    def tile_stats(self):
        """(flux evaluations per cell, most rounds of a sub-tile, tiles, cells per tile) of the device layout"""
        a, b, c_, d = C.c_double(), C.c_int32(), C.c_int32(), C.c_int32()
        self.c.lib.check(self.c.lib.dll.adfvm_tile_stats(self.c.ctx, C.byref(a), C.byref(b), C.byref(c_), C.byref(d)))
        return a.value, b.value, c_.value, d.value

    def tile_rounds(self):
        """(rounds summed over all sub-tiles, sub-tiles, early tiles): rounds/sub-tiles = 4.0 on a regular hex block"""
        a, b, e = C.c_int64(), C.c_int64(), C.c_int32()
        self.c.lib.check(self.c.lib.dll.adfvm_tile_rounds(self.c.ctx, C.byref(a), C.byref(b), C.byref(e)))
        return a.value, b.value, e.value

    def tile_halo_stats(self):
        """(largest tile halo, kernel variant, tiles per halo-size bin of 32 slots)"""
        a, b = C.c_int32(), C.c_int32()
        h = (C.c_int32 * 32)()
        self.c.lib.check(self.c.lib.dll.adfvm_tile_halo_stats(self.c.ctx, C.byref(a), C.byref(b), h, 32))
        return a.value, b.value, list(h)

    def _prepare(self, inputs, options):
        opts = dict(self.defaultOptions)
        for k in options:
            if k not in opts:
                raise TypeError("unknown option %r" % k)
        opts.update(options)
        P = self.c.parse(inputs)
        if not self.c.static_loaded:
            self.c.load_static(P)
        elif opts["replace_static"]:
            self.c.load_replaceable(P)
        if P["sizes"] != self.c.sizes:
            raise ValueError("mesh size constants changed between calls")
        return P, opts

    def __call__(self, *inputs, **options):
        c = self.c
        P, opts = self._prepare(inputs, options)
        if P["rest"]:
            raise NotImplementedError("extraArgs are not supported on the hot path")
        C_ = c.sizes[2]
        rho, rhoU, rhoE = P["state"]
        c._arr(rho, (1,), "rho", C_); c._arr(rhoU, (3,), "rhoU", C_); c._arr(rhoE, (1,), "rhoE", C_)
        dt = c._arr(P["dt"], (1,), "dt", 1)
        flags = (L.RETURN_REUSABLE if opts["return_reusable"] else 0) | (L.REPLACE_REUSABLE if opts["replace_reusable"] else 0)
        outs = [None, None, None]
        if opts["return_reusable"]:
            outs = [c.pool.empty((C_, 1), c.dtype), c.pool.empty((C_, 3), c.dtype), c.pool.empty((C_, 1), c.dtype)]
        dtc, obj = np.zeros((1, 1), c.dtype), np.zeros((1, 1), c.dtype)
        cached = opts["replace_reusable"] and c.cache_select((rho, rhoU, rhoE))
        rc = c.lib.dll.adfvm_primal(c.ctx, None if cached else _ptr(rho), None if cached else _ptr(rhoU), None if cached else _ptr(rhoE),
                                    float(dt[0, 0]), flags, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(dtc), _ptr(obj))
        c.check_traced()
        c.lib.check(rc)
        if opts["return_reusable"]:
            c.cache_register(outs)
        return (outs[0], outs[1], outs[2], dtc, obj)

    def init_fields(self, *inputs):
        """Drop-in for `solver.mapBoundary` (Function('init', ...), adFVM/density.py:64-80): inputs = (rho, rhoU, rhoE) + mesh
        arrays + constants + patch triples + BC arrays; returns the conservative fields with ghost rows, [nCells][d]."""
        c = self.c
        ins = list(inputs)
        if not c.static_loaded:           # `init` has no dt / source arguments: complete the list to the layout of `primal`
            C_ = int(ins[3 + 15 + 2])
            k = 3 + 15 + 8 + 3 * len(c.sorted)
            full = ins[:3] + [np.zeros((1, 1), c.dtype)] + ins[3:k] + \
                [np.zeros((C_, 1), c.dtype), np.zeros((C_, 3), c.dtype), np.zeros((C_, 1), c.dtype)] + ins[k:]
            if (c.spec.get("objective") or {}).get("kind") == "traced":
                raise RuntimeError("call primal once before init (the traced objective binds the inputs of `primal`)")
            c.load_static(c.parse(full))
        C_, N = c.sizes[2], c.sizes[0]
        rho, rhoU, rhoE = ins[:3]
        c._arr(rho, (1,), "rho", C_); c._arr(rhoU, (3,), "rhoU", C_); c._arr(rhoE, (1,), "rhoE", C_)
        outs = [np.empty((N, 1), c.dtype), np.empty((N, 3), c.dtype), np.empty((N, 1), c.dtype)]
        c.lib.check(c.lib.dll.adfvm_init_fields(c.ctx, _ptr(rho), _ptr(rhoU), _ptr(rhoE), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])))
        return tuple(outs)

    def grad(self):
        """Counterpart of `primal.map.grad()` (adpy/adpy/variable.py:322-343): the reverse-mode function of
        this step for parameters='source' (apps/adjoint.py:94-126)."""
        return AdjointFunction(self)

    # resident stepping used by the benchmark and by long runs between report steps
    def step_resident(self, dt):
        self.c.lib.check(self.c.lib.dll.adfvm_primal_step_resident(self.c.ctx, float(dt)))

    def set_state(self, *inputs):
        """upload mesh/BC/source (first time) and the state from a `primal` positional list, without stepping"""
        c = self.c
        P, _ = self._prepare(inputs, {})
        C_ = c.sizes[2]
        rho, rhoU, rhoE = P["state"]
        c._arr(rho, (1,), "rho", C_); c._arr(rhoU, (3,), "rhoU", C_); c._arr(rhoE, (1,), "rhoE", C_)
        c.lib.check(c.lib.dll.adfvm_set_state(c.ctx, _ptr(rho), _ptr(rhoU), _ptr(rhoE)))

    def state(self):
        """(rho, rhoU, rhoE) currently resident (what the next call without replace_reusable continues from)"""
        c = self.c
        C_ = c.sizes[2]
        outs = [c.pool.empty((C_, 1), c.dtype), c.pool.empty((C_, 3), c.dtype), c.pool.empty((C_, 1), c.dtype)]
        c.lib.check(c.lib.dll.adfvm_get_state(c.ctx, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])))
        return tuple(outs)

    def run_block(self, dts):
        """len(dts) resident primal steps; the state at the start of every step stays on the device for
        AdjointFunction.run_block (replaces the host-side `solutions` list of Solver.run(mode='forward'),
        adFVM/solver.py:376-382). Returns (dtc[k], objective[k])."""
        n = len(dts)
        dt = (C.c_double * max(n, 1))(*[float(x) for x in dts])
        dtc, obj = (C.c_double * max(n, 1))(), (C.c_double * max(n, 1))()
        self.c.lib.check(self.c.lib.dll.adfvm_primal_block(self.c.ctx, n, dt, dtc, obj))
        return np.array(dtc[:n]), np.array(obj[:n])

    def dtc_obj(self):
        a, b = C.c_double(), C.c_double()
        self.c.lib.check(self.c.lib.dll.adfvm_get_dtc_obj(self.c.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def dtc_global(self):
        """dtc of the last step maximised over the ranks (a collective; NCCL inside the library)"""
        a = C.c_double()
        self.c.lib.check(self.c.lib.dll.adfvm_get_dtc_global(self.c.ctx, C.byref(a)))
        return a.value

    def next_dt(self, dt, CFL, stepFactor=1.2, remaining=float("inf")):
        """the adaptive time step of Solver.run (adFVM/solver.py:365): min(2*CFL/dtc over all ranks, dt*stepFactor, endTime-t)"""
        return min(2. * CFL / self.dtc_global(), dt * stepFactor, remaining)

    def sync(self):
        self.c.lib.check(self.c.lib.dll.adfvm_sync(self.c.ctx))

    def kernel_timing(self, enable):
        self.c.lib.check(self.c.lib.dll.adfvm_kernel_timing(self.c.ctx, int(bool(enable))))

    def kernel_report(self):
        """{kernel: (launches, total_ms)} accumulated since kernel_timing(True)"""
        buf = C.create_string_buffer(8192)
        self.c.lib.check(self.c.lib.dll.adfvm_kernel_report(self.c.ctx, buf, 8192))
        out = {}
        for line in buf.value.decode().strip().split("\n"):
            if line:
                k, n, ms = line.split()
                out[k] = (int(n), float(ms))
        return out

    @property
    def state_cache_hits(self):
        """calls that started from a cached device copy instead of an upload"""
        return int(self.c.lib.dll.adfvm_state_cache_hits(self.c.ctx))

    @property
    def launches(self):
        return int(self.c.lib.dll.adfvm_launch_count(self.c.ctx))

    @property
    def graph_replays(self):
        """steps served by replaying a captured CUDA graph"""
        return int(self.c.lib.dll.adfvm_graph_replays(self.c.ctx))

    @property
    def device_bytes(self):
        return int(self.c.lib.dll.adfvm_device_bytes(self.c.ctx))


class AdjointFunction:
    """Drop-in for `adjoint.map` (Function('primal_grad', ...), apps/adjoint.py:124-126) with
    parameters='source': inputs = primal inputs + [rhoa, rhoUa, rhoEa, dtca, obja] + [scaling];
    returns (rhoa', rhoUa', rhoEa', dJ/dS_rho, dJ/dS_rhoU, dJ/dS_rhoE); the three gradients are None unless
    return_static (static accumulators, zeroed on zero_static)."""
    name = "primal_grad"
    defaultOptions = PrimalFunction.defaultOptions

    def __init__(self, primal, viscosity=None, rtol=0., maxit=0):
        self.primal = primal
        self.c = primal.c
        # viscosity: None, or the adjParams[1] of the case file ('abarbanel', 'turkel', 'uniform'): this object then is
        # `adjoint.viscousMap` (Function('primal_grad_viscous'), apps/adjoint.py:127-141)
        if viscosity not in L.VISC:
            raise NotImplementedError("adjoint viscosity type %r (supported: abarbanel, turkel, uniform; 'entropy_hughes' is "
                                      "generated Mathematica code + a generalised eigenproblem in the reference)" % (viscosity,))
        self.viscosity, self.visc_rtol, self.visc_maxit = viscosity, float(rtol), int(maxit)

    def viscous(self, viscosity, rtol=0., maxit=0):
        """drop-in for `adjoint.viscousMap`: same inputs and outputs as this function; the returned adjoint fields have had one
        step of the adjoint artificial viscosity applied (M_2norm of the step's start state scaled by the `scaling` input,
        adFVM/postpro.py:491-720)"""
        return AdjointFunction(self.primal, viscosity, rtol, maxit)

    def adjoint_viscosity(self, rho, rhoU, rhoE, scaling):
        """diagnostic (write_M_2norm of apps/adjoint.py): M_2norm [nCells,1] of a state for this function's viscosity type"""
        c = self.c
        C_ = c.sizes[2]
        c._arr(rho, (1,), "rho", C_); c._arr(rhoU, (3,), "rhoU", C_); c._arr(rhoE, (1,), "rhoE", C_)
        c.lib.check(c.lib.dll.adfvm_set_adjoint_viscosity(c.ctx, L.VISC[self.viscosity], float(scaling), self.visc_rtol, self.visc_maxit))
        M = np.zeros((c.sizes[0], 1), c.dtype)
        c.lib.check(c.lib.dll.adfvm_get_adjoint_viscosity(c.ctx, _ptr(rho), _ptr(rhoU), _ptr(rhoE), _ptr(M)))
        return M

    @property
    def viscosity_iterations(self):
        return int(self.c.lib.dll.adfvm_viscosity_iterations(self.c.ctx))

    def viscous_resident(self, dt, scaling):
        """the smoothing applied to the resident adjoint fields (after step_resident / run_block)"""
        c = self.c
        c.lib.check(c.lib.dll.adfvm_set_adjoint_viscosity(c.ctx, L.VISC[self.viscosity], float(scaling), self.visc_rtol, self.visc_maxit))
        c.lib.check(c.lib.dll.adfvm_adjoint_viscous_resident(c.ctx, float(dt)))

    def __call__(self, *inputs, **options):
        c = self.c
        P, opts = self.primal._prepare(inputs, options)
        rest = P["rest"]
        if len(rest) != 6:
            raise TypeError("primal_grad expects primal inputs + 5 adjoint seeds + scaling (got %d extra)" % len(rest))
        C_ = c.sizes[2]
        rho, rhoU, rhoE = P["state"]
        c._arr(rho, (1,), "rho", C_); c._arr(rhoU, (3,), "rhoU", C_); c._arr(rhoE, (1,), "rhoE", C_)
        dt = c._arr(P["dt"], (1,), "dt", 1)
        ra, rUa, rEa = rest[0:3]
        c._arr(ra, (1,), "rhoa", C_); c._arr(rUa, (3,), "rhoUa", C_); c._arr(rEa, (1,), "rhoEa", C_)
        dtca = float(c._arr(rest[3], (1,), "dtca", 1)[0, 0]); obja = float(c._arr(rest[4], (1,), "obja", 1)[0, 0])
        flags = (L.RETURN_STATIC if opts["return_static"] else 0) | (L.ZERO_STATIC if opts["zero_static"] else 0)
        if self.viscosity is not None:
            scaling = float(c._arr(rest[5], (1,), "scaling", 1)[0, 0])
            c.lib.check(c.lib.dll.adfvm_set_adjoint_viscosity(c.ctx, L.VISC[self.viscosity], scaling, self.visc_rtol, self.visc_maxit))
            flags |= L.VISCOUS
        outs = [c.pool.empty((C_, 1), c.dtype), c.pool.empty((C_, 3), c.dtype), c.pool.empty((C_, 1), c.dtype)]
        grads = [None, None, None]
        if c.param_mesh:                           # the ten metric arrays: read after the call (adfvm_get_mesh_grad)
            grads = [None, None, None]
        elif c.param_bc is not None:               # one BC input array is the parameter block
            grads = [c.pool.empty(c.param_bc, c.dtype) if opts["return_static"] else None, None, None]
        elif opts["return_static"]:
            grads = [c.pool.empty((C_, 1), c.dtype), c.pool.empty((C_, 3), c.dtype), c.pool.empty((C_, 1), c.dtype)]
        cached = c.cache_select((rho, rhoU, rhoE))
        rc = c.lib.dll.adfvm_primal_grad(
            c.ctx, None if cached else _ptr(rho), None if cached else _ptr(rhoU), None if cached else _ptr(rhoE), float(dt[0, 0]),
            _ptr(ra), _ptr(rUa), _ptr(rEa), dtca, obja, flags,
            _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2]))
        c.check_traced()
        c.lib.check(rc)
        if c.param_mesh:
            if not opts["return_static"]:
                if opts["zero_static"]:
                    self.mesh_gradients(zero_static=True)
                return tuple(outs + [None] * 10)
            return tuple(outs + self.mesh_gradients(zero_static=opts["zero_static"]))
        if c.param_bc is not None:
            return tuple(outs + grads[:1])
        return tuple(outs + grads)

    def mesh_gradients(self, zero_static=False):
        """parameters='mesh': the accumulated gradients of the ten metric arrays, in Mesh.gradFields order
        (adFVM/mesh.py:27-31), shaped like the inputs"""
        c = self.c
        F, Cn, Fi = c.sizes[1], c.sizes[2], c.sizes[3]
        shapes = [(F, 1), (F, 1), (Fi, 1), (F, 1), (F, 1), (F, 3), (F, 3), (F, 2), (F, 2, 3), (Cn, 1)]
        arrs = [np.zeros(sh, c.dtype) for sh in shapes]
        c.lib.check(c.lib.dll.adfvm_get_mesh_grad(c.ctx, *[_ptr(a) for a in arrs], int(zero_static)))
        traced = getattr(c, "traced", None)           # a traced objective differentiates its own use of the metrics (torch side)
        extra = traced.take_mesh_grad(zero_static) if traced is not None else None
        if extra is not None:
            for a, e in zip(arrs, extra):
                a += e.reshape(a.shape)
        return arrs

    def step_resident(self, dt, obja=1.0, chain=True):
        self.c.lib.check(self.c.lib.dll.adfvm_adjoint_step_resident(self.c.ctx, float(dt), float(obja), int(chain)))

    def set_fields(self, rhoa, rhoUa, rhoEa):
        """upload the adjoint fields the next run_block / step_resident(chain=True) starts from"""
        c = self.c
        C_ = c.sizes[2]
        c._arr(rhoa, (1,), "rhoa", C_); c._arr(rhoUa, (3,), "rhoUa", C_); c._arr(rhoEa, (1,), "rhoEa", C_)
        c.lib.check(c.lib.dll.adfvm_set_adjoint(c.ctx, _ptr(rhoa), _ptr(rhoUa), _ptr(rhoEa)))

    def run_block(self, dts, obja=1.0, scaling=0.0):
        """reverse sweep over the block stored by PrimalFunction.run_block(dts) (the step loop of Adjoint.run,
        apps/adjoint.py:250-291, without the host round trips): adjoint fields and source-term gradient stay resident"""
        n = len(dts)
        dt = (C.c_double * max(n, 1))(*[float(x) for x in dts])
        c = self.c
        if self.viscosity is None:
            c.lib.check(c.lib.dll.adfvm_adjoint_block(c.ctx, n, dt, float(obja)))
        else:                                       # this object is `viscousMap`: the smoothing follows every step of the block
            c.lib.check(c.lib.dll.adfvm_set_adjoint_viscosity(c.ctx, L.VISC[self.viscosity], float(scaling), self.visc_rtol, self.visc_maxit))
            c.lib.check(c.lib.dll.adfvm_adjoint_block_viscous(c.ctx, n, dt, float(obja)))

    def fields(self, return_static=True, zero_static=False):
        """(rhoa, rhoUa, rhoEa[, dJ/dS_rho, dJ/dS_rhoU, dJ/dS_rhoE]) currently resident"""
        c = self.c
        C_ = c.sizes[2]
        outs = [c.pool.empty((C_, 1), c.dtype), c.pool.empty((C_, 3), c.dtype), c.pool.empty((C_, 1), c.dtype)]
        grads = [None, None, None]
        if return_static:
            grads = [c.pool.empty((C_, 1), c.dtype), c.pool.empty((C_, 3), c.dtype), c.pool.empty((C_, 1), c.dtype)]
        c.lib.check(c.lib.dll.adfvm_get_adjoint(c.ctx, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]),
                                                _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2]), int(zero_static)))
        return tuple(outs + grads)


def spec_from_solver(primal, objective):
    """Build the static spec from a reference `RCF` object (adFVM/density.py) after `readFields`.
    `objective` is one of the supported objective dicts (the reference takes arbitrary adpy-DSL code here)."""
    mesh = primal.mesh
    patches = []
    for pid in list(mesh.sortedPatches) + list(mesh.remotePatches):
        p = mesh.boundary[pid]
        d = {"name": pid, "type": p["type"], "startFace": int(p["startFace"]), "nFaces": int(p["nFaces"]),
             "cellStartFace": int(p["cellStartFace"])}
        for k in ("neighbourPatch", "myProcNo", "neighbProcNo", "referPatch"):
            if k in p:
                d[k] = p[k] if isinstance(p[k], str) else int(p[k])
        if pid in mesh.remotePatches:
            d["tag"] = int(mesh.getProcessorPatchInfo(pid)[2])
        patches.append(d)
    bcs = {phi.name: {pid: {"type": bc.__class__.__name__, "keys": list(bc.keys)} for pid, bc in phi.phi.BC.items()}
           for phi in primal.fields}
    T = np.array([250., 300., 400.])
    muv = np.asarray(primal.mu(T), np.float64) * np.ones(3)
    if np.allclose(muv, 1.4792e-06 * T ** 1.5 / (T + 116.), rtol=1e-14, atol=0):
        mu = {"law": "sutherland"}
    elif np.all(muv == muv[0]):
        mu = {"law": "constant", "value": float(muv[0])}
    else:
        raise NotImplementedError("viscosity law must be constant or Sutherland")
    return {"Cp": primal.Cp, "gamma": primal.gamma, "Pr": primal.Pr, "mu": mu,
            "riemannSolver": primal.riemannSolver.__name__,
            "boundaryRiemannSolver": primal.boundaryRiemannSolver.__name__,
            "timeIntegrator": primal.timeIntegrator, "patches": patches, "BCs": bcs,
            "sortedPatches": list(mesh.sortedPatches), "objective": objective}
