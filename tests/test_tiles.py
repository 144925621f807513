"""Host-side tile plan (fvm_tiles.h) through the C ABI of the CPU simulator: structured blocks must decompose into
exact 4x4x8 tiles / 4x4x2 sub-tiles whatever the block size (4.0 flux evaluations per cell, 4 full rounds per sub-tile,
160 halo slots), including sizes that a plain halving of the tile count cuts raggedly (3*32, 5*32 planes) and the
sampled lattice detection used on large meshes."""
import os

import numpy as np
import pytest

from adfvm_b200 import cases, function


def _plan(hostsim, n):
    case = cases.periodic_box(n)
    f = function.PrimalFunction(case.spec, np.float64, lib=hostsim)
    f.c.load_static(f.c.parse(case.inputs()))
    return f.tile_stats(), f.tile_rounds(), f.tile_halo_stats()


@pytest.mark.parametrize("n", [(16, 16, 16), (24, 24, 24), (40, 24, 8), (12, 20, 24)])
def test_structured_blocks_tile_regularly(hostsim, n):
    (evals, max_rounds, ntiles, T), (rounds, subtiles, early), (max_halo, variant, _) = _plan(hostsim, n)
    assert T == 128 and variant == 0
    assert evals == 4.0 and max_rounds == 4 and rounds == 4 * subtiles
    assert max_halo == 160
    assert early == ntiles                      # no processor patches: every tile is "early"


def test_sampled_lattice_detection(hostsim, monkeypatch):
    """large meshes detect their lattice from a scattered sample of the cells (a regular stride aliases with the
    lattice and misses planes); force the sampled path on a small mesh"""
    monkeypatch.setenv("ADFVM_LATTICE_SAMPLE", "3000")
    (evals, max_rounds, ntiles, T), (rounds, subtiles, _), (max_halo, variant, _) = _plan(hostsim, (24, 24, 24))
    assert evals == 4.0 and max_rounds == 4 and rounds == 4 * subtiles and max_halo == 160


def test_irregular_sizes_still_work(hostsim):
    """extents that are not multiples of 4: ragged tiles, more rounds, same results (checked by the golden tests)"""
    (evals, max_rounds, ntiles, T), (rounds, subtiles, _), _ = _plan(hostsim, (10, 9, 7))
    assert 3.0 < evals < 6.0 and max_rounds >= 4
