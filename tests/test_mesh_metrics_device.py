"""Mesh metric build through the C ABI (adfvm_mesh_metrics, csrc/fvm_metrics.h) against the numpy
restatement of adFVM/cpp/cmesh.cpp (adfvm_b200/metrics.py, itself pinned to the reference's cmesh outputs recorded in
the golden fixtures by tests/test_metrics.py): warped periodic box (cyclic ghost centres), graded walled channel, half
annulus, box with a hole. CPU simulator here; the device in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from adfvm_b200 import hexmesh
from adfvm_b200.metrics import build_mesh, GRAD_FIELDS, INT_FIELDS


def polys():
    lo, hi = (0., 0., 0.), (1., 1., 1.)
    yield "warped periodic box", hexmesh.box_mesh((6, 5, 4), lo, hi, warp=hexmesh.sine_warp(0.03, lo, hi))
    yield "graded channel", hexmesh.box_mesh((8, 6, 3), lo, (2., 1., .5), grading=(1.0, 0.4, 1.0), patches=[
        ("inlet", "patch", ["x-"], {}), ("outlet", "patch", ["x+"], {}), ("floor", "symmetryPlane", ["y-"], {}),
        ("lid", "patch", ["y+"], {}), ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}), ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])

    def warp(p):
        r = 0.5e-3 * 16. ** p[:, 0]
        return np.stack([r * np.cos(np.pi * p[:, 1]), r * np.sin(np.pi * p[:, 1]), p[:, 2]], axis=1)
    yield "half annulus", hexmesh.box_mesh((12, 16, 1), lo, (1., 1., 2e-4), warp=warp, patches=[
        ("cylinder", "patch", ["x-"], {}), ("far", "patch", ["x+"], {}), ("axis", "symmetryPlane", ["y-", "y+"], {}),
        ("z1", "cyclic", ["z-"], {"neighbourPatch": "z2"}), ("z2", "cyclic", ["z+"], {"neighbourPatch": "z1"})])
    K, J, I = np.meshgrid(np.arange(1), np.arange(10), np.arange(30), indexing="ij")
    yield "forward step", hexmesh.masked_box_mesh((30, 10, 1), (0., 0., -0.05), (3., 1., 0.05), ~((I >= 6) & (J < 2)), [
        ("inlet", "patch", ["x-"], {}), ("outlet", "patch", ["x+"], {}), ("bottom", "symmetryPlane", ["y-"], {}),
        ("top", "symmetryPlane", ["y+"], {}), ("defaultFaces", "empty", ["z-", "z+"], {})])


def compare(lib, precision=np.float64, tol=1e-12, **dev):
    n = 0
    for name, poly in polys():
        ref = build_mesh(poly)
        got = build_mesh(poly, device=dict(lib=lib, precision=precision, **dev))
        for a in GRAD_FIELDS + ["faceCentres", "cellCentres"]:
            x, y = np.asarray(getattr(got, a), np.float64), np.asarray(getattr(ref, a), np.float64)
            assert x.shape == y.shape, (name, a)
            assert np.abs(x - y).max() <= tol * np.abs(y).max(), (name, a, np.abs(x - y).max() / np.abs(y).max())
        for a in INT_FIELDS:
            assert np.array_equal(getattr(got, a), getattr(ref, a)), (name, a)
        assert got.getScalar() == ref.getScalar()
        n += 1
    assert n == 4


def test_metrics_through_the_c_abi(hostsim):
    compare(hostsim)


def test_metrics_fp32(hostsim):
    compare(hostsim, np.float32, 2e-4)      # quadratic weights are differences of nearby points: fp32 keeps ~4 digits
