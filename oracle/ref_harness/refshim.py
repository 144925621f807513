"""TEST INFRASTRUCTURE ONLY — makes the *unmodified* reference (/root/reference) importable and
runnable in this container so golden vectors can be generated from it (SURVEY §8(c) recipe).

Nothing here is imported by the product package. Nothing is copied from the reference: its
sources are compiled where they lie, outputs go to oracle/_ref/ (git-ignored) and to scratch
case directories. Only usable where /root/reference exists (the build container).

What the shim does (each was required for a green run of the reference on python 3.12 /
numpy 2 / no MPI / no LAPACK):
  1. sys.path <- reference root + adpy submodule
  2. fake single-rank `mpi4py.MPI` module
  3. fake `ar` module (noise column of objective.txt only)
  4. numpy-2 compatibility: np.fromstring(bytes) and np.product, ndarray.tostring users patched
  5. builds `cmesh` (and fp32 `cmesh_gpu`) natives from the reference sources with g++
  6. overrides config.get_compiler_args: gcc/g++ instead of ccache mpicc, stub mpi.h, stub LAPACK
  7. optional recorder around adpy.variable.Function.__call__ (dumps every map call's
     positional inputs + outputs)
  8. optional DROP-IN mode (env ADFVM_DROPIN_LIB + ADFVM_DROPIN_OBJECTIVE): every call of the reference's compiled
     `primal` / `primal_grad` functions is served by adfvm_b200.function.PrimalFunction / AdjointFunction built with
     spec_from_solver from the reference's own RCF object - the reference's drivers, time loop, checkpointing and file
     output run unmodified around them (tests/test_dropin_reference.py)
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import subprocess
import sys
import sysconfig
import types

import numpy as np

REF = os.environ.get("ADFVM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
STUBS = os.path.join(HERE, "stubs")


# ----------------------------------------------------------------------------- fake modules
def _install_fake_mpi():
    class _Comm:
        def Get_size(self): return 1
        def Get_rank(self): return 0
        def gather(self, x, root=0): return [x]
        def bcast(self, x, root=0): return x
        def Bcast(self, x, root=0): return None
        def Barrier(self): return None
        def scatter(self, x, root=0): return x[0]
        def Allreduce(self, a, b, op=None): b[...] = a
        def Reduce(self, a, b, op=None, root=0): b[...] = a
        def allreduce(self, x, op=None): return x
        def Exscan(self, a, b, op=None): b[...] = 0
        def Scan(self, a, b, op=None): b[...] = a
        def Abort(self, code=1): raise SystemExit(code)
        def Isend(self, *a, **k): raise RuntimeError("single-rank MPI shim: Isend")
        def Irecv(self, *a, **k): raise RuntimeError("single-rank MPI shim: Irecv")

    class _Request:
        @staticmethod
        def Waitall(reqs): return None

    MPI = types.ModuleType("mpi4py.MPI")
    MPI.COMM_WORLD = _Comm()
    MPI.MAX, MPI.MIN, MPI.SUM, MPI.MINLOC = "MAX", "MIN", "SUM", "MINLOC"
    MPI.Request = _Request
    MPI.Get_processor_name = lambda: "localhost"
    pkg = types.ModuleType("mpi4py")
    pkg.MPI = MPI
    sys.modules["mpi4py"] = pkg
    sys.modules["mpi4py.MPI"] = MPI


def _install_fake_ar():
    ar = types.ModuleType("ar")

    class _Sel:
        mu_sigma = np.array([0.])
    ar.arsel = lambda x: _Sel()
    sys.modules["ar"] = ar


def _build_cfuncs():
    """adFVM/compat/cfuncs.pyx (Cython: intersectPlane of the vane objective, ...) compiled from where it lies: cython writes
    the C++ to a scratch directory (language level 2: the file has python-2 integer division, cfuncs.pyx:77), g++ the module
    into oracle/_ref/"""
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "cfuncs" + sysconfig.get_config_var("EXT_SUFFIX"))
    pyx = os.path.join(REF, "adFVM", "compat", "cfuncs.pyx")
    if os.path.exists(so) and os.path.getmtime(so) > os.path.getmtime(pyx):
        return so
    import tempfile
    tmp = tempfile.mkdtemp(prefix="cfuncs_")
    cpp = os.path.join(tmp, "cfuncs.cpp")
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-2", pyx, "-o", cpp])
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-w", "-fopenmp", "-I" + sysconfig.get_paths()["include"],
                           "-I" + np.get_include(), cpp, "-o", so])
    return so


def _install_fake_cfuncs():
    """adFVM/compat/cfuncs.pyx is a Cython module (cut-plane geometry for the vane objective); density.py
    imports postpro.py which imports it. The reference's own module is used when it has been built into oracle/_ref
    (_build_cfuncs); otherwise a stub whose functions abort (they are not on the hot path)."""
    so = os.path.join(OUT, "cfuncs" + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(so):
        cf = _load_ext("cfuncs", so, "adFVM.compat.cfuncs")
        pkg = types.ModuleType("adFVM.compat")
        pkg.__path__ = []
        pkg.cfuncs = cf
        for n in ("intersectPlane", "reduceAbsMin", "selectMultipleRange", "reduceSum"):
            setattr(pkg, n, getattr(cf, n))
        sys.modules["adFVM.compat"] = pkg
        return

    def _unavailable(*a, **k):
        raise RuntimeError("adFVM.compat.cfuncs is stubbed in the oracle harness")
    cf = types.ModuleType("adFVM.compat.cfuncs")
    for n in ("intersectPlane", "reduceAbsMin", "selectMultipleRange", "reduceSum", "decompose"):
        setattr(cf, n, _unavailable)
    pkg = types.ModuleType("adFVM.compat")
    pkg.__path__ = []
    pkg.cfuncs = cf
    for n in ("intersectPlane", "reduceAbsMin", "selectMultipleRange", "reduceSum", "decompose"):
        setattr(pkg, n, _unavailable)
    sys.modules["adFVM.compat"] = pkg
    sys.modules["adFVM.compat.cfuncs"] = cf


def _numpy_compat():
    _orig = getattr(np, "fromstring")

    def fromstring(s, dtype=float, count=-1, sep=""):
        if isinstance(s, (bytes, bytearray, memoryview)) and sep == "":
            return np.frombuffer(s, dtype=dtype, count=count).copy()
        return _orig(s, dtype=dtype, count=count, sep=sep)
    np.fromstring = fromstring
    if not hasattr(np, "product"):
        np.product = np.prod


# ----------------------------------------------------------------------------- native builds
def _build_cmesh(fp32=False):
    os.makedirs(OUT, exist_ok=True)
    name = "cmesh_gpu" if fp32 else "cmesh"
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, name + ext)
    cpp = os.path.join(REF, "adFVM", "cpp")
    srcs = [os.path.join(cpp, f) for f in ("mesh.cpp", "cmesh.cpp")] + \
           [os.path.join(REF, "adpy", "adpy", "cpp", "interface.cpp")]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(s) for s in srcs):
        return so
    inc = [os.path.join(cpp, "include"), os.path.join(REF, "adpy", "adpy", "cpp", "include"),
           sysconfig.get_paths()["include"], np.get_include(), STUBS]
    cmd = ["g++", "-std=c++11", "-O3", "-fPIC", "-shared", "-fopenmp", "-UNDEBUG", "-w",
           ] + (["-DCPU_FLOAT32"] if fp32 else []) + \
          ["-I" + i for i in inc] + srcs + ["-o", so]
    subprocess.check_call(cmd)
    return so


def _load_ext(modname, path, alias):
    loader = importlib.machinery.ExtensionFileLoader(modname, path)
    spec = importlib.util.spec_from_loader(modname, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[alias] = mod
    return mod


# ----------------------------------------------------------------------------- recorder
RECORD = []          # list of (name, inputs, options, outputs)
RECORD_ON = [False]


DROPIN = {"lib": os.environ.get("ADFVM_DROPIN_LIB"), "objective": os.environ.get("ADFVM_DROPIN_OBJECTIVE"),
          "rcf": None, "f": None, "calls": 0}


def _dropin_function():
    """PrimalFunction for the RCF object the driver compiled (built at the first call: the BCs exist by then)"""
    if DROPIN["f"] is None:
        import json
        root = os.path.dirname(os.path.dirname(HERE))
        if root not in sys.path:
            sys.path.insert(0, root)
        from adfvm_b200 import function as b200, _lib
        from adFVM import config
        spec = b200.spec_from_solver(DROPIN["rcf"], json.loads(DROPIN["objective"]))
        DROPIN["f"] = b200.PrimalFunction(spec, config.precision, lib=_lib.Lib(DROPIN["lib"]))
        DROPIN["g"] = DROPIN["f"].grad()
    return DROPIN["f"]


def _install_dropin():
    import atexit
    from adFVM.density import RCF
    orig = RCF.compileSolver
    atexit.register(lambda: print("[dropin] served %d calls through %s" % (DROPIN["calls"], DROPIN["lib"]), flush=True))

    def compileSolver(self):
        orig(self)
        DROPIN["rcf"] = self
    RCF.compileSolver = compileSolver


def _install_recorder():
    from adpy.variable import Function
    orig = Function.__call__

    def call(self, *args, **kwargs):
        if DROPIN["lib"] and self.name in ("primal", "primal_grad"):
            f = _dropin_function()
            DROPIN["calls"] += 1
            out = (f if self.name == "primal" else DROPIN["g"])(*args, **kwargs)
        else:
            out = orig(self, *args, **kwargs)
        if RECORD_ON[0]:
            def cp(x):
                return np.array(x, copy=True) if isinstance(x, np.ndarray) else x
            RECORD.append((self.name, [cp(a) for a in args], dict(kwargs),
                           [cp(o) for o in (out if isinstance(out, (tuple, list)) else [out])]))
        return out
    Function.__call__ = call


# ----------------------------------------------------------------------------- entry point
def install(fp32=False, argv=None, overlay=None):
    """Call BEFORE importing adFVM. argv replaces sys.argv (adFVM.config parses it at import).
    overlay: a directory put in FRONT of the reference's on sys.path (the `adpy` drop-in overlay of adfvm_b200/dropin: the
    equivalent of prepending it to PYTHONPATH); the recorder / compile overrides of this shim are then not installed."""
    if argv is not None:
        sys.argv = list(argv)
    for p in (os.path.join(REF, "adpy"), REF, os.path.join(REF, "apps")):
        if p not in sys.path:
            sys.path.insert(0, p)
    if overlay:
        sys.path.insert(0, overlay)
    _install_fake_mpi()
    _install_fake_ar()
    _install_fake_cfuncs()
    _numpy_compat()

    so = _build_cmesh(fp32=False)
    import adFVM  # noqa: F401  (package init is empty)
    import adFVM.cpp  # noqa: F401
    if fp32:
        so32 = _build_cmesh(fp32=True)
        mod = _load_ext("cmesh_gpu", so32, "adFVM.cpp.cmesh_gpu")
        sys.modules["adFVM.cpp.cmesh"] = mod
        adFVM.cpp.cmesh = mod
        adFVM.cpp.cmesh_gpu = mod
    else:
        mod = _load_ext("cmesh", so, "adFVM.cpp.cmesh")
        adFVM.cpp.cmesh = mod

    from adFVM import config
    import adpy.config
    if fp32:
        config.precision = adpy.config.precision = np.float32
        config.SMALL, config.VSMALL, config.LARGE = 1e-9, 1e-30, 1e30

    cppDir = os.path.join(REF, "adFVM", "cpp")

    def get_compiler_args():
        # LAPACK (only the adjoint-viscosity path calls it): abort stubs, or forwarders to the library named by ADFVM_LAPACK_SO
        lapack = "lapack_fwd.cpp" if os.environ.get("ADFVM_LAPACK_SO") else "lapack_stub.cpp"
        return {"compiler": "gcc", "linker": "g++", "libs": ["dl"] if lapack == "lapack_fwd.cpp" else [],
                "incdirs": [os.path.join(cppDir, "include"), STUBS],
                "libdirs": [],
                "sources": [os.path.join(cppDir, x) for x in
                            ["external.cpp", "mesh.cpp", "parallel.cpp", "scaling.cpp"]] +
                           [os.path.join(STUBS, lapack)],
                "extra_compile_args": ["-w"] + (["-DCPU_FLOAT32"] if fp32 else [])}
    config.get_compiler_args = get_compiler_args

    # ndarray.tostring was removed in numpy 2: the reference's writers use it (mesh.py:1052)
    import adFVM.mesh as rmesh
    import adFVM.field as rfield

    def writeField(handle, field, dtype, initial):
        handle.write((initial + " nonuniform List<" + dtype + ">\n").encode())
        handle.write(("{0}\n(".format(len(field))).encode())
        handle.write(np.ascontiguousarray(field, np.float64).tobytes())
        handle.write(")\n;\n".encode())
    rmesh.writeField = writeField
    rfield.writeField = writeField

    # py3 str/bytes: readFoam decodes non-binary boundary entries to str (field.py:253) but
    # extractField matches them with bytes regexes (mesh.py:1032-1033) -> TypeError for any
    # `uniform ...` boundary value. Re-encode before calling the reference's own function.
    import adFVM.BCs as rbcs
    import adFVM.solver as rsolver
    _extract = rmesh.extractField

    def extractField(data, size, dimensions):
        if isinstance(data, str):
            data = data.encode()
        return _extract(data, size, dimensions)
    for mod in (rmesh, rfield, rbcs, rsolver):
        if hasattr(mod, "extractField"):
            mod.extractField = extractField
    if overlay:
        return config
    _install_recorder()
    if DROPIN["lib"]:
        _install_dropin()
    return config
