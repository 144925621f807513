"""TEST INFRASTRUCTURE ONLY — run one of the reference's own drivers (apps/problem.py or
apps/adjoint.py, unmodified) under the compatibility shim and record every compiled-function
call (`primal`, `init`, `primal_grad`): positional inputs, options, outputs.

usage: python run_ref.py {problem|adjoint} RECORD.npz [--fp32] -- <driver argv...>
"""
import json
import os
import runpy
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402


def describe(primal):
    """Static description of the compiled problem, read off the reference's own objects."""
    mesh = primal.mesh
    patches = []
    for pid in mesh.sortedPatches + mesh.remotePatches:
        p = mesh.boundary[pid]
        d = {"name": pid, "type": p["type"], "startFace": int(p["startFace"]), "nFaces": int(p["nFaces"]),
             "cellStartFace": int(p["cellStartFace"])}
        for k in ("neighbourPatch", "myProcNo", "neighbProcNo", "referPatch"):
            if k in p:
                d[k] = p[k] if isinstance(p[k], str) else int(p[k])
        patches.append(d)
    bcs = {}
    for phi in primal.fields:
        bcs[phi.name] = {pid: {"type": bc.__class__.__name__, "keys": list(bc.keys)}
                         for pid, bc in phi.phi.BC.items()}
    T = np.array([250., 300., 400.])
    muv = np.asarray(primal.mu(T), np.float64) * np.ones(3)
    if np.allclose(muv, 1.4792e-06 * T ** 1.5 / (T + 116.), rtol=1e-14, atol=0):
        mu = {"law": "sutherland"}
    else:
        assert np.all(muv == muv[0]), "unsupported viscosity law"
        mu = {"law": "constant", "value": float(muv[0])}
    return {"Cp": primal.Cp, "gamma": primal.gamma, "Pr": primal.Pr, "mu": mu,
            "riemannSolver": primal.riemannSolver.__name__,
            "boundaryRiemannSolver": primal.boundaryRiemannSolver.__name__,
            "timeIntegrator": primal.timeIntegrator, "patches": patches, "BCs": bcs,
            "sortedPatches": list(mesh.sortedPatches)}


def main():
    app = sys.argv[1]
    record = sys.argv[2]
    rest = sys.argv[3:]
    fp32 = "--fp32" in rest[:rest.index("--")] if "--" in rest else False
    argv = rest[rest.index("--") + 1:]
    script = os.path.join(refshim.REF, "apps", app + ".py")
    refshim.install(fp32=fp32, argv=[script] + argv)
    refshim.RECORD_ON[0] = True
    g = runpy.run_path(script, run_name="__main__")
    primal = g["primal"]
    calls = refshim.RECORD
    out = {"n_calls": np.array(len(calls))}
    meta = {"calls": [], "spec": describe(primal)}
    static_seen = None
    for ci, (name, inputs, options, outputs) in enumerate(calls):
        kinds = []
        for ii, a in enumerate(inputs):
            if isinstance(a, np.ndarray):
                # mesh / BC / source arrays repeat verbatim in every call: store once
                key0 = "c0_i%d" % ii
                if ci > 0 and key0 in out and out[key0].shape == a.shape and np.array_equal(out[key0], a):
                    kinds.append("same")
                else:
                    out["c%d_i%d" % (ci, ii)] = a
                    kinds.append("array")
            else:
                kinds.append(int(a))
        outs = []
        for oi, o in enumerate(outputs):
            if isinstance(o, np.ndarray):
                out["c%d_o%d" % (ci, oi)] = o
                outs.append("array")
            else:
                outs.append(None)
        meta["calls"].append({"name": name, "inputs": kinds, "options": {k: bool(v) for k, v in options.items()},
                              "outputs": outs})
    mesh = primal.mesh
    for a in ("cellCentres", "faceCentres", "points", "faces"):
        out["mesh_" + a] = np.asarray(getattr(mesh, a))
    np.savez_compressed(record, **out)
    with open(record.replace(".npz", ".json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
