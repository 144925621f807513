"""TEST INFRASTRUCTURE ONLY - golden vectors of the reference's adjoint artificial viscosity.

The UNMODIFIED reference computes them: `adFVM/postpro.py:491-690 computeAdjointViscosity` (the traced kernels, the default
ghost fill and the face interpolation) is compiled by the reference's own adpy code generator next to `primal`, with
`adFVM/cpp/scaling.cpp:84-106 Function_get_max_eigenvalue` calling LAPACK's dsyev (resolved from scipy's bundled OpenBLAS
through stubs/lapack_fwd.cpp - the image has no system LAPACK). Nothing of the reference is edited or copied: this harness
only asks its code generator for one more Function per viscosity type, which is what `apps/adjoint.py:127-141` does when a
case file sets `adjParams = [scaling, type, None]`.

The implicit diffusion that follows (`Function_apply_adjoint_viscosity`, scaling.cpp:134-160) needs PETSc or the reference's
CUDA build (`matop_petsc.cpp`, `matop_cuda.cpp`): neither exists here, the reference then copies its input through
(scaling.cpp:153-158). That part is pinned by the oracle's exact sparse solve of the system `matop_petsc.cpp:288-372` assembles.

usage: python oracle/ref_harness/gen_viscosity.py [case ...]      -> tests/golden/visc_<case>.npz
Each fixture: the state (the case's initial fields as the reference's `primal` received them), `scaling`, and per type
M_2norm [nCells,1] (ghost rows filled) and DT [nFaces,1].
"""
import glob
import os
import runpy
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
TYPES = ("abarbanel", "turkel", "uniform")
SCALING = 1e-3


def lapack_so():
    import scipy
    libs = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*"))
    assert libs, "scipy's bundled OpenBLAS not found"
    return libs[0]


def child(casefile, out):
    os.environ["ADFVM_LAPACK_SO"] = lapack_so()
    import refshim
    script = os.path.join(refshim.REF, "apps", "problem.py")
    refshim.install(argv=[script, casefile, "-c"])
    refshim.RECORD_ON[0] = True
    from adFVM.density import RCF
    VISC = {}
    orig = RCF.compileSolver

    def compileSolver(self):          # apps/adjoint.py:127-141 asks for the same Function (there followed by the diffusion solve)
        orig(self)
        from adFVM.postpro import computeAdjointViscosity
        from adpy.variable import Variable, Function
        for vt in TYPES:
            scaling = Variable((1, 1))
            fields = list(self.map._inputs[:3])
            M_2norm, DT = computeAdjointViscosity(*([self, vt] + fields + [scaling]))
            VISC[vt] = Function("visc_" + vt, list(self.map._inputs) + [scaling], [M_2norm, DT])
    RCF.compileSolver = compileSolver
    runpy.run_path(script, run_name="__main__")
    calls = [c for c in refshim.RECORD if c[0] == "primal"]
    inputs = calls[0][1]
    res = {"rho": inputs[0], "rhoU": inputs[1], "rhoE": inputs[2], "scaling": np.array(SCALING)}
    refshim.RECORD_ON[0] = False
    for vt in TYPES:
        M, DT = VISC[vt](*(list(inputs) + [np.array([[SCALING]], inputs[0].dtype)]))
        res["M_2norm_" + vt], res["DT_" + vt] = np.array(M), np.array(DT)
        print(vt, "M_2norm", M.shape, float(M.min()), float(M.max()), "DT", DT.shape, flush=True)
    np.savez_compressed(out, **res)


def main():
    import subprocess
    import gen_golden
    names = sys.argv[1:] or ["box_walls", "box_cyclic", "cyl2d"]
    for name in names:
        c, case, casefile = gen_golden.write_case(name, "visc_" + name)
        out = os.path.join(gen_golden.GOLDEN, "visc_%s.npz" % name)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", casefile, out], cwd=case)
        print("wrote", out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3])
    else:
        main()
