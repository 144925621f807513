"""TEST / BASELINE INFRASTRUCTURE ONLY — loader for the reference's own compiled step functions (oracle/_ref/refgraph_*.so,
built by oracle/ref_harness/build_ref_graph.py from the unmodified reference's generated code and runtime sources).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` legs may import this module; the
product package never does. The module needs nothing from /root/reference at run time: the .so is self-contained
(single-rank MPI stub linked in), and `initialize(rank, mesh)` (adpy/adpy/cpp/module/graph.cpp:16-30 ->
adFVM/cpp/external.cpp:21-30 -> adFVM/cpp/mesh.cpp:5-49) only reads plain attributes of the mesh object it is given.

    g = RefGraph("box_cyclic")            # raises FileNotFoundError when the .so was not built / did not travel
    g.initialize(mesh)                    # adfvm_b200.metrics.MeshData (or anything with the same attributes)
    rho, rhoU, rhoE, dtc, obj = g.primal(*inputs, replace_reusable=True)
    rhoa, rhoUa, rhoEa, gS0, gS1, gS2 = g.primal_grad(*adjoint_inputs)

Calling convention = adFVM/solver.py:312-317 and apps/adjoint.py:272-280; options = adpy/adpy/variable.py:282-287.
"""
import importlib.machinery
import importlib.util
import os
import sysconfig
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_OPTIONS = {"return_static": True, "zero_static": False, "replace_static": False,
                   "return_reusable": True, "replace_reusable": False}     # adpy/adpy/variable.py:282-287


def module_path(name, fp32=False, native=False):
    return os.path.join(HERE, "_ref", "refgraph_%s_%s%s%s" % (name, "f32" if fp32 else "f64", "_native" if native else "",
                                                               sysconfig.get_config_var("EXT_SUFFIX")))


def _native_usable(path):
    """the -march=native variant only on a host whose CPU has every feature of the build host"""
    import json
    try:
        need = set(json.load(open(path.split(".")[0] + ".json"))["cpu_flags"])
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return need <= set(line.split(":", 1)[1].split())
    except Exception:
        pass
    return False


def available(name="box_cyclic", fp32=False):
    return os.path.exists(module_path(name, fp32))


class RefGraph:
    def __init__(self, name="box_cyclic", fp32=False):
        path = module_path(name, fp32)
        self.variant = "x86-64-v3"
        npath = module_path(name, fp32, native=True)
        if os.path.exists(npath) and _native_usable(npath) and not os.environ.get("ADFVM_REFGRAPH_PORTABLE"):
            path, self.variant = npath, "march=native of the build host"
        if not os.path.exists(path):
            raise FileNotFoundError("%s not built (python oracle/ref_harness/build_ref_graph.py %s, needs /root/reference)" % (path, name))
        modname = os.path.basename(path).split(".")[0]
        loader = importlib.machinery.ExtensionFileLoader(modname, path)
        spec = importlib.util.spec_from_loader(modname, loader)
        self.mod = importlib.util.module_from_spec(spec)
        loader.exec_module(self.mod)
        self.dtype = np.float32 if fp32 else np.float64
        self._mesh = None

    def initialize(self, mesh, rank=0):
        """mesh: the attributes adFVM/cpp/mesh.cpp:5-49 reads"""
        m = types.SimpleNamespace()
        for a in ("nCells", "nFaces", "nInternalFaces", "nInternalCells", "nGhostCells"):
            setattr(m, a, int(getattr(mesh, a)))
        m.nBoundaryFaces = m.nFaces - m.nInternalFaces
        m.nLocalCells = int(getattr(mesh, "nLocalCells", m.nCells))
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        sc = lambda a: np.ascontiguousarray(a, self.dtype)
        # geometry inputs of cmesh.build: not read by the step functions; placeholders of the right rank when absent
        m.faces = i32(getattr(mesh, "faces5", np.zeros((m.nFaces, 5), np.int32)))
        m.points = sc(getattr(mesh, "points", np.zeros((1, 3))))
        m.owner, m.neighbour = i32(mesh.owner), i32(mesh.neighbour)
        m.boundary = {str(k): {str(kk): (vv if isinstance(vv, str) else int(vv)) for kk, vv in v.items() if isinstance(vv, (str, int, np.integer))}
                      for k, v in mesh.boundary.items()}
        remote = [k for k, v in mesh.boundary.items() if v["type"] in ("processor", "processorCyclic")]
        if remote:
            raise NotImplementedError("the prebuilt reference module is single-rank (MPI stub)")
        m.nLocalPatches, m.nRemotePatches, m.tags = len(mesh.boundary), 0, {}
        m.cellNeighboursMatOp = i32(mesh.cellNeighbours)
        m.areas, m.deltas, m.volumes = sc(mesh.areas), sc(mesh.deltas), sc(mesh.volumes)
        m.cellFaces = i32(mesh.cellFaces)
        self._mesh = m            # the C++ side keeps borrowed array pointers
        # the module's Py_AtExit hook (graph.cpp:38-41 -> external.cpp:44-47) drops its reference after the interpreter has
        # been finalised: keep one more so that nothing is deallocated there
        import ctypes
        ctypes.pythonapi.Py_IncRef(ctypes.py_object(m))
        self.mod.initialize(rank, m)

    def _call(self, name, args, kwargs):
        opts = dict(DEFAULT_OPTIONS)
        opts.update(kwargs)
        return getattr(self.mod, name)(*args, **opts)

    def primal(self, *args, **kwargs):
        return self._call("primal", args, kwargs)

    def primal_grad(self, *args, **kwargs):
        return self._call("primal_grad", args, kwargs)


# ---- one mesh per process. The reference's runtime keeps its static buffers (gradient accumulators, adpy/adpy/cpp/include/
# common.hpp:312-327 `shared_acquire`) in a process-wide table keyed by variable id and sized by the FIRST mesh it sees, as a
# solver process of the reference only ever has one mesh. Checks that visit several meshes therefore run each RefGraph in
# its own child process.
def _serve(conn, name, fp32):
    try:
        g = RefGraph(name, fp32)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)                           # "Initializing C++ interface"
        conn.send(("ok", g.variant))
        while True:
            msg = conn.recv()
            if msg[0] == "close":
                break
            try:
                if msg[0] == "initialize":
                    g.initialize(msg[1]); conn.send(("ok", None))
                else:
                    conn.send(("ok", getattr(g, msg[0])(*msg[1], **msg[2])))
            except Exception as e:      # noqa: BLE001
                conn.send(("error", repr(e)))
    except Exception as e:              # noqa: BLE001
        conn.send(("error", repr(e)))
    os._exit(0)                                       # skip the module's exit hook


class IsolatedRefGraph:
    """RefGraph in a child process (same methods); close() or garbage collection ends the child"""

    def __init__(self, name="box_cyclic", fp32=False):
        import multiprocessing as mp
        if not available(name, fp32):
            raise FileNotFoundError(module_path(name, fp32))
        ctx = mp.get_context("spawn")
        self.conn, child = ctx.Pipe()
        self.proc = ctx.Process(target=_serve, args=(child, name, fp32), daemon=True)
        self.proc.start()
        self.variant = self._reply()

    def _reply(self):
        status, val = self.conn.recv()
        if status != "ok":
            raise RuntimeError("reference process: %s" % val)
        return val

    def initialize(self, mesh):
        self.conn.send(("initialize", mesh)); self._reply()

    def primal(self, *args, **kwargs):
        self.conn.send(("primal", args, kwargs)); return self._reply()

    def primal_grad(self, *args, **kwargs):
        self.conn.send(("primal_grad", args, kwargs)); return self._reply()

    def close(self):
        if self.proc is not None:
            try:
                self.conn.send(("close",))
            except Exception:
                pass
            self.proc.join(timeout=10)
            if self.proc.is_alive():
                self.proc.kill()
            self.proc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
