"""The drop-in boundary with NO edit of the reference and NO declared objective: the reference's own drivers (apps/problem.py,
apps/adjoint.py, unmodified) run with the `adpy` overlay of adfvm_b200/dropin in front of the reference's adpy on the module
path (= PYTHONPATH only). The case file - including its adpy-DSL objective (drag of templates/cylinder_test.py:9-36, the vane
design objective of adFVM/objectives/vane.py with its extraArgs and intermediate mpi_allreduce) - is traced by the reference's
own front-end; `primal`, `primal_grad` and `init` are served by adfvm_b200 (here: the CPU simulator of the device code), the
objective by adfvm_b200.adpy_objective from that trace. objective.txt must equal the stock reference's (golden fixtures,
recorded from the unmodified reference): `orig` and `adjoint` to 1e-9 relative, `perturb` to 1e-11 of the objective it is a
difference of. Needs /root/reference (the build container); elsewhere the tests skip."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ADFVM_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")

CASES = [("cyl2d", ("problem", "perturb", "adjoint")),               # traced drag objective, Lax-Friedrichs boundary solver
         ("channel_vane", ("problem", "adjoint")),                   # vane objective: extraArgs + allreduce inside the objective
         ("box_walls_mesh", ("problem", "adjoint"))]                 # parameters = 'mesh': the objective's own metric gradient
if os.environ.get("ADFVM_SLOW_TESTS"):
    CASES += [("box_walls_bcpt", ("problem", "adjoint")), ("step2d", ("problem", "perturb", "adjoint")), ("tube", ("problem", "adjoint")),
              ("box_upt", ("problem", "adjoint"))]


@pytest.mark.parametrize("name,runs", CASES)
def test_reference_drivers_with_overlay(name, runs, hostsim):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_harness"))
    import gen_golden
    c, case, casefile = gen_golden.write_case(name, name + "_overlay")
    env = dict(os.environ, ADFVM_DROPIN_LIB=hostsim.path)
    runner = os.path.join(ROOT, "oracle", "ref_harness", "run_overlay.py")
    for r in runs:
        app = "adjoint" if r == "adjoint" else "problem"
        argv = [casefile] + (["perturb"] if r == "perturb" else [])
        out = subprocess.run([sys.executable, runner, app, "--"] + argv, cwd=case, env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
        served = [ln for ln in out.stdout.split("\n") if ln.startswith("[overlay]")]
        assert served and "dropin/adpy/__init__.py" in served[-1] and "PrimalFunction" in served[-1], served
        assert not os.path.isdir(os.path.join(case, "gencode"))          # nothing was generated or compiled
    got = {ln.split()[0]: float(ln.split()[3]) for ln in open(os.path.join(case, "objective.txt")).read().strip().split("\n")}
    with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as f:
        ref = {ln.split()[0]: float(ln.split()[3]) for ln in json.load(f)["objective_txt"]}
    assert abs(got["orig"] - ref["orig"]) <= 1e-9 * abs(ref["orig"])
    assert abs(got["adjoint"] - ref["adjoint"]) <= 1e-9 * abs(ref["adjoint"])
    if "perturb" in runs:
        assert abs(got["perturb"] - ref["perturb"]) <= 1e-11 * abs(ref["orig"])
