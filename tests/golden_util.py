"""Loader for tests/golden/*.npz fixtures (recorded calls of the unmodified reference;
generator: oracle/ref_harness/gen_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name):
        self.name = name
        with open(os.path.join(GOLDEN, name + ".json")) as f:
            self.meta = json.load(f)
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.spec = self.meta["spec"]

    def calls(self, run, fname=None):
        """yield (index, name, inputs, options, outputs) for one recorded driver run"""
        for ci, c in enumerate(self.meta["runs"][run]):
            if fname is not None and c["name"] != fname:
                continue
            inputs = []
            for ii, kind in enumerate(c["inputs"]):
                if kind == "array":
                    inputs.append(self.z["%s__c%d_i%d" % (run, ci, ii)])
                elif kind == "same":
                    inputs.append(self.z["%s__c0_i%d" % (run, ii)])
                else:
                    inputs.append(int(kind))
            outputs = [self.z["%s__c%d_o%d" % (run, ci, oi)] if k == "array" else None
                       for oi, k in enumerate(c["outputs"])]
            yield ci, c["name"], inputs, c["options"], outputs

    def mesh_array(self, run, name):
        return self.z["%s__mesh_%s" % (run, name)]


def available():
    """fixtures recorded call by call (the anchors of BASELINE.md section 5 are listed by anchors())"""
    return sorted(f[:-5] for f in os.listdir(GOLDEN) if f.endswith(".json") and not f.startswith("anchor_"))


def anchors():
    return sorted(f[:-5] for f in os.listdir(GOLDEN) if f.endswith(".json") and f.startswith("anchor_"))


class Anchor:
    """objective.txt, time series and strided samples of the final fields of one anchor run of the unmodified reference
    (oracle/ref_harness/gen_golden.py generate_anchor)"""

    def __init__(self, name):
        with open(os.path.join(GOLDEN, name + ".json")) as f:
            self.meta = json.load(f)
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.spec, self.stride = self.meta["spec"], int(self.meta["stride"])
        self.cf = self.meta["casefile"]
        self.objective = {ln.split()[0]: float(ln.split()[3]) for ln in self.meta["objective_txt"]}

    def check_fields(self, prefix, names, arrs, tol, scales=None):
        """compare arrays with the stored strided sample and norms of `<prefix>_<name>`; error relative to the group's max"""
        scales = scales or [1.0] * len(arrs)
        num = den = 0.0
        ssq_den = max(float(self.z["%s_%s_norms" % (prefix, n)][1].max()) * s * s for n, s in zip(names, scales))
        for n, a, s in zip(names, arrs, scales):
            a = np.asarray(a, np.float64).reshape(len(a), -1)
            ref = self.z["%s_%s_sample" % (prefix, n)]
            num = max(num, float(np.abs(a[::self.stride] - ref).max()) * s)
            den = max(den, float(self.z["%s_%s_norms" % (prefix, n)][2].max()) * s)
            nr = self.z["%s_%s_norms" % (prefix, n)]
            ssq = (a * a).sum(axis=0)
            assert np.all(np.abs(ssq - nr[1]) * s * s <= 10 * tol * max(ssq_den, 1e-300)), (prefix, n, "sum of squares")
        err = num / max(den, 1e-300)
        assert err < tol, (prefix, err)
        return err


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def relerr_cols(a, b, floor=1e-3):
    """column by column: each component against its OWN largest value, but not below `floor` times the array's largest - a
    component that is small everywhere (rhoU_z of a 2-D case) is the sum of terms of the size of the large ones, whose rounding
    sets its error floor; below that a relative error has no meaning. With floor = 1e-3 a component 1000 times smaller than the
    largest is still checked to the full tolerance of its own size."""
    a = np.asarray(a, np.float64).reshape(len(a), -1); b = np.asarray(b, np.float64).reshape(len(b), -1)
    top = max(float(np.abs(b).max()), 1e-300)
    return max(float(np.abs(a[:, k] - b[:, k]).max()) / max(float(np.abs(b[:, k]).max()), floor * top) for k in range(b.shape[1]))


def group_relerr(arrs, refs, scales=None):
    """Largest error over a group of arrays that share one physical unit after scaling
    (e.g. adjoints of rho, rhoU, rhoE scaled by typical rho, rhoU, rhoE), relative to the group's max."""
    scales = scales or [1.0] * len(arrs)
    num = max(float(np.abs(np.asarray(a, np.float64) - np.asarray(r, np.float64)).max()) * s
              for a, r, s in zip(arrs, refs, scales))
    den = max(float(np.abs(np.asarray(r, np.float64)).max()) * s for r, s in zip(refs, scales))
    return num / max(den, 1e-300)


def state_scales(inputs):
    return [float(np.abs(inputs[i]).max()) for i in range(3)]
