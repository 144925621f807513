"""TEST / BASELINE INFRASTRUCTURE ONLY — build the reference's OWN compiled step functions for a given problem class
into oracle/_ref/ so that they can be timed (bench.py --impl reference, cpu_baseline kind "reference") and used as
the parity checker at sizes the small golden fixtures do not reach, on boxes where /root/reference does not exist.

What is built: the CPython extension the reference itself generates and loads as `graph_N`
(adpy/adpy/variable.py:545-590): the unmodified reference drivers (apps/adjoint.py through the shim of refshim.py) trace
the case file and write gencode/{code.cpp, kernel.cpp, kernel.hpp}; those generated sources and the reference's own
runtime sources (adFVM/cpp/{external,mesh,parallel,scaling}.cpp, adpy/adpy/cpp/{interface.cpp,module/graph.cpp}) are
compiled from where they lie with the reference's flags (-O3 -march=native, adpy/adpy/compile.py:37-58). Because the binary
travels to another host, two variants are built: `*_native` (with a side-car JSON of the build host's CPU flags; the
loader takes it only on a host that has all of them) and a portable `-march=x86-64-v3` one. Nothing of the reference is copied into the repository:
the output is a .so under oracle/_ref/ (git-ignored).

The module exports `initialize(rank, mesh)`, `init`, `primal`, `primal_grad` with the positional calling convention of
adFVM/solver.py:312-317 / apps/adjoint.py:272-280. The generated code takes every mesh array and size as an argument, so
one module serves any mesh with the same patches, boundary conditions, objective and parameter block.

usage: python oracle/ref_harness/build_ref_graph.py [box_cyclic ...]
"""
import os
import subprocess
import sys
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF = os.environ.get("ADFVM_REFERENCE", "/root/reference")


def module_path(name, fp32=False, native=False):
    return os.path.join(OUT, "refgraph_%s_%s%s%s" % (name, "f32" if fp32 else "f64", "_native" if native else "",
                                                      sysconfig.get_config_var("EXT_SUFFIX")))


def cpu_flags():
    for line in open("/proc/cpuinfo"):
        if line.startswith("flags"):
            return sorted(set(line.split(":", 1)[1].split()))
    return []


def build(name="box_cyclic", fp32=False, force=False):
    import gen_golden
    so = module_path(name, fp32)
    if os.path.exists(so) and not force:
        return so
    os.makedirs(OUT, exist_ok=True)
    tag = "refgraph_" + name + ("_f32" if fp32 else "")
    c, case, casefile = gen_golden.write_case(name, tag)
    runner = os.path.join(HERE, "run_ref.py")
    flag = ["--fp32"] if fp32 else []
    # the reference traces the case file and generates + compiles + runs its own module (few steps of a tiny mesh)
    # (the adjoint driver reads the checkpoints the primal driver writes, so that one runs first)
    for app in ("problem", "adjoint"):
        subprocess.check_call([sys.executable, runner, app, os.path.join(case, "rec_%s.npz" % app)] + flag + ["--", casefile, "-c"], cwd=case)
    gen = os.path.join(case, "gencode")
    mod = os.path.basename(so).split(".")[0]
    cpp = os.path.join(REF, "adFVM", "cpp")
    acpp = os.path.join(REF, "adpy", "adpy", "cpp")
    stubs = os.path.join(HERE, "stubs")
    srcs = [os.path.join(cpp, f) for f in ("external.cpp", "mesh.cpp", "parallel.cpp", "scaling.cpp")] + \
           [os.path.join(acpp, "interface.cpp"), os.path.join(acpp, "module", "graph.cpp"), os.path.join(stubs, "lapack_stub.cpp"),
            os.path.join(gen, "kernel.cpp"), os.path.join(gen, "code.cpp")]
    inc = [os.path.join(cpp, "include"), stubs, sysconfig.get_paths()["include"], np.get_include(), gen, os.path.join(acpp, "include")]
    import json
    for native in (False, True):
        target = module_path(name, fp32, native)
        mod = os.path.basename(target).split(".")[0]
        flags = ["-w", "-std=c++11", "-O3", "-DMODULE=" + mod, "-fPIC", "-march=native" if native else "-march=x86-64-v3"] + \
                (["-DCPU_FLOAT32"] if fp32 else [])
        objs, procs = [], []
        for s in srcs:
            o = os.path.join(gen, "rg%d_" % native + os.path.basename(s).replace(".cpp", ".o"))
            objs.append(o)
            procs.append(subprocess.Popen(["g++"] + flags + ["-I" + i for i in inc] + ["-c", s, "-o", o]))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("compile failed")
        subprocess.check_call(["g++", "-shared"] + objs + ["-o", target])
        if native:
            with open(target.split(".")[0] + ".json", "w") as f:
                json.dump({"march": "native", "cpu_flags": cpu_flags()}, f)
    return so


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["box_cyclic"]
    for n in names:
        print(build(n, fp32="--fp32" in sys.argv, force="--force" in sys.argv))
