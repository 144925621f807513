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
    return sorted(f[:-5] for f in os.listdir(GOLDEN) if f.endswith(".json"))


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


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
