"""Rank-local pieces of a block-decomposed periodic hex box, in the reference's decomposed-case conventions
(`processorN/constant/polyMesh/boundary`, reference adFVM/mesh.py:238-249, 746-756, 784-805):

* every rank owns an (nx,ny,nz) block of a (nx*px, ny*py, nz*pz) periodic box (weak scaling: fixed work per rank);
* a block side whose direction is not split (p == 1) stays a local `cyclic` pair, exactly as on one rank;
* a side in a split direction becomes a `processor` patch (interior cut) or a `processorCyclic` patch (the cut
  coincides with the periodic wrap), named `procBoundary<i>to<j>[through<cyc>]` like OpenFOAM's decomposePar,
  listed after all physical patches so that remote ghost rows are the tail of the cell arrays, with
  `myProcNo/neighbProcNo` and a `tag` shared by both sides; face i of a patch matches face i of the peer's patch.

The halo itself (pack -> NCCL send/recv -> unpack, reverse + add for the adjoint) lives in the native library
(adfvm_comm_init, replacing adFVM/cpp/parallel.cpp:31-209); `attach_comm` distributes the communicator id over
torch.distributed, which is plumbing only.
"""
from __future__ import annotations

import numpy as np

from . import cases
from .metrics import uniform_box


def factor3(world):
    """(px, py, pz) with px >= py >= pz, as cubic as possible: 1->(1,1,1) 2->(2,1,1) 4->(2,2,1) 8->(2,2,2)"""
    best = None
    for px in range(1, world + 1):
        if world % px:
            continue
        for py in range(1, world // px + 1):
            if (world // px) % py:
                continue
            pz = world // px // py
            if px >= py >= pz:
                key = (px - pz, px)
                if best is None or key < best[0]:
                    best = (key, (px, py, pz))
    return best[1]


def rank_coords(rank, p):
    return rank % p[0], (rank // p[0]) % p[1], rank // (p[0] * p[1])


def coords_rank(c, p):
    return (c[0] % p[0]) + p[0] * ((c[1] % p[1]) + p[1] * (c[2] % p[2]))


def _side_tag(dim, plus, me, peer):
    # identical on both ends of a connection: my '+' side meets the peer's '-' side
    return 2 * dim + (1 if plus == (me < peer) else 0)


def periodic_box_rank(n, rank, world, dtype=np.float64, p=None, dt=None):
    """Case of `rank` out of `world` (see module docstring). n: cells per side per rank (int or 3-tuple)."""
    if isinstance(n, int):
        n = (n, n, n)
    p = p or factor3(world)
    assert p[0] * p[1] * p[2] == world
    me = rank_coords(rank, p)
    L = tuple(1.0 for _ in range(3))                      # block edge length; global box = p[d] * L[d]
    lo = tuple(me[d] * L[d] for d in range(3))
    hi = tuple(lo[d] + L[d] for d in range(3))
    phys, proc = [], []
    names = "xyz"
    for d in range(3):
        c1, c2 = names[d] + "1", names[d] + "2"
        if p[d] == 1:
            phys.append((c1, "cyclic", [names[d] + "-"], {"neighbourPatch": c2}))
            phys.append((c2, "cyclic", [names[d] + "+"], {"neighbourPatch": c1}))
            continue
        for plus in (False, True):
            nb = list(me); nb[d] += 1 if plus else -1
            wrap = nb[d] < 0 or nb[d] >= p[d]
            peer = coords_rank(nb, p)
            extra = {"myProcNo": rank, "neighbProcNo": peer, "tag": _side_tag(d, plus, rank, peer)}
            name = "procBoundary%dto%d" % (rank, peer)
            ptype = "processor"
            if wrap:
                ptype = "processorCyclic"
                extra["referPatch"] = c2 if plus else c1
                name += "through" + extra["referPatch"]
            proc.append((name, ptype, [names[d] + ("+" if plus else "-")], extra))
    mesh = uniform_box(n, lo, hi, phys + proc)            # ghost centres of processor faces = the peer block's cells
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    U, T, pr = cases.smooth_state(cc)                     # period 1 in every direction: periodic on the global box too
    bcs = {f: {pid: {"type": "cyclic", "keys": []} for pid in mesh.sortedPatches} for f in ("U", "T", "p")}
    spec = cases._spec(mesh, bcs, {"kind": "cell_TV"})
    if dt is None:
        dt = 1e-6 * 48. / max(n)
    mid = tuple(0.5 * p[d] * L[d] for d in range(3))
    return cases.Case(mesh, spec, cases.conservative(U, T, pr), cases.gaussian_source(cc, mid), {}, dt, dtype)


def global_box(n, world, dtype=np.float64, p=None, dt=None):
    """The undecomposed mesh of the same problem (single rank), for decomposition-invariance checks
    (the reference's own criterion, tests/test_parallel.py:63-81)."""
    if isinstance(n, int):
        n = (n, n, n)
    p = p or factor3(world)
    N = tuple(n[d] * p[d] for d in range(3))
    hi = tuple(float(p[d]) for d in range(3))
    mesh = uniform_box(N, (0., 0., 0.), hi)
    C = mesh.nInternalCells
    cc = mesh.cellCentres[:C]
    U, T, pr = cases.smooth_state(cc)
    bcs = {f: {pid: {"type": "cyclic", "keys": []} for pid in mesh.sortedPatches} for f in ("U", "T", "p")}
    spec = cases._spec(mesh, bcs, {"kind": "cell_TV"})
    if dt is None:
        dt = 1e-6 * 48. / max(n)
    mid = tuple(0.5 * hi[d] for d in range(3))
    return cases.Case(mesh, spec, cases.conservative(U, T, pr), cases.gaussian_source(cc, mid), {}, dt, dtype)


def global_cell_ids(n, rank, world, p=None):
    """global cell index (in `global_box` numbering) of every cell of `rank`'s block"""
    if isinstance(n, int):
        n = (n, n, n)
    p = p or factor3(world)
    me = rank_coords(rank, p)
    K, J, I = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
    gi, gj, gk = I.ravel() + me[0] * n[0], J.ravel() + me[1] * n[1], K.ravel() + me[2] * n[2]
    return gi + n[0] * p[0] * (gj + n[1] * p[1] * gk)


def attach_comm(f, rank, world, group=None):
    """Create the native halo communicator of PrimalFunction `f`: rank 0 makes the id, torch.distributed
    (already initialised by the caller: nccl on GPUs) broadcasts it."""
    import ctypes
    import torch
    import torch.distributed as dist
    lib = f.c.lib
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        lib.check(lib.dll.adfvm_comm_unique_id(buf))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0, group=group)
    f.c.attach_comm(bytes(t.cpu().tolist()), rank, world)
