"""TEST INFRASTRUCTURE ONLY - run one of the reference's own drivers (apps/problem.py or apps/adjoint.py, unmodified) with the
`adpy` drop-in overlay of adfvm_b200/dropin in front of the reference's adpy on sys.path (what `PYTHONPATH=<repo>/adfvm_b200/dropin:
<reference>/adpy:<reference>` does), plus the environment shims this container needs to import the reference at all (no MPI,
numpy 2: refshim.py items 2-5). No recorder, no compiler: every `primal` / `primal_grad` / `init` call of the driver is served
by adfvm_b200 (library chosen by ADFVM_DROPIN_LIB, default the CUDA product library).

usage: python run_overlay.py {problem|adjoint} [--fp32] -- <driver argv...>
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import refshim  # noqa: E402


def main():
    app = sys.argv[1]
    rest = sys.argv[2:]
    fp32 = "--fp32" in rest[:rest.index("--")]
    argv = rest[rest.index("--") + 1:]
    script = os.path.join(refshim.REF, "apps", app + ".py")
    refshim.install(fp32=fp32, argv=[script] + argv, overlay=os.path.join(ROOT, "adfvm_b200", "dropin"))
    runpy.run_path(script, run_name="__main__")
    import adpy
    mod = adpy._STATE["module"]
    print("[overlay] adpy = %s; primal served by %s (%d kernel launches); viscous calls %d" %
          (adpy.__file__, type(mod.primal_f).__name__, mod.primal_f.launches, mod.viscous_calls), flush=True)


if __name__ == "__main__":
    main()
