"""ORACLE - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy / scipy.sparse) of adFVM's adjoint artificial-viscosity stabilisation:
the `primal_grad_viscous` function `apps/adjoint.py:127-141` builds when a case file sets `adjParams = [scaling, type, None]`.

Parity status:
  * M_2norm / DT (everything up to the diffusion solve): PINNED against outputs of the unmodified reference
    (tests/golden/visc_*.npz, made by oracle/ref_harness/gen_viscosity.py: the reference's own traced kernels +
    LAPACK dsyev), tests/test_viscosity.py.
  * the implicit diffusion solve: PARITY UNPINNED. The reference solves it with PETSc (GMRES + hypre, `matop_petsc.cpp:374-440`,
    default rtol 1e-5) or with 1000 Jacobi sweeps in its CUDA build (`matop_cuda.cpp:201-231`); neither library exists in
    this image and without them the reference copies its input through (`scaling.cpp:153-158`). `apply_adjoint_viscosity`
    below assembles the system exactly as `matop_petsc.cpp:288-372` / `matop_cuda.cpp:160-178` do and solves it directly
    (scipy.sparse.linalg.spsolve), i.e. it is the fixed point both reference solvers iterate towards.

Paths relative to /root/reference. Supported types: 'abarbanel', 'turkel', 'uniform' ('entropy_hughes' is generated
Mathematica code + a generalised eigenproblem, not restated).
"""
from __future__ import annotations

import numpy as np

from . import adfvm_oracle as O

TYPES = ("abarbanel", "turkel", "uniform")
UREF, TREF, PREF = 33., 300., 1e5           # adFVM/density.py:57-59


def _np(t):
    return t.detach().numpy() if hasattr(t, "detach") else np.asarray(t)


def _fields_with_ghosts(P):
    """U, T, p [nCells,*] of the state in P: primitive + the ghost fill `computeAdjointViscosity` traces
    (adFVM/postpro.py:527-533: boundaryInit / boundary / boundaryEnd = updateGhostCells + characteristic BCs)."""
    import torch
    C, nC = P.nInternalCells, P.nCells
    Ui, Ti, pi = O.primitive(P, *P.state)
    pad = lambda x: torch.cat([x, torch.zeros((nC - C,) + x.shape[1:], dtype=x.dtype)], dim=0)
    U, T, p = pad(Ui), pad(Ti), pad(pi)
    U = O.update_ghost(P, U, "U"); T = O.update_ghost(P, T, "T"); p = O.update_ghost(P, p, "p")
    U, T, p = O.characteristic_boundary(P, U, T, p)
    return _np(U), _np(T), _np(p)


def viscosity_gradients(P, U, T, p):
    """`gradients` kernel of adFVM/postpro.py:499-524 over internal faces (both sides), coupled patches (owner side,
    central face values) and the other patches (owner side, face value = ghost value): op.grad / op.div of the face values
    (adFVM/op.py:12-43). Returns gradU [C,3,3] (gradU[i][j] = dU_i/dx_j), divU [C,1], gradp [C,3], gradc [C,3]."""
    C, Fi = P.nInternalCells, P.nInternalFaces
    g, Rg = P.gamma, P.Cp - P.Cv
    own, nei = _np(P.owner), _np(P.neighbour)
    A, N, w = _np(P.areas), _np(P.normals), _np(P.weights)
    VL, VR = _np(P.volumesL), _np(P.volumesR)
    gradU, divU = np.zeros((C, 3, 3)), np.zeros((C, 1))
    gradp, gradc = np.zeros((C, 3)), np.zeros((C, 3))

    def faces(sl, neighbour, boundary):
        o, n = own[sl], nei[sl]
        if boundary:                                               # postpro.py:504-508
            UF, pF, TF = U[n], p[n], T[n]
            cF = np.sqrt(g * TF * Rg)
        else:                                                      # postpro.py:509-515 (interp.central, interp.py:15-18)
            f = w[sl]
            UF = U[o] * f + U[n] * (-f + 1)
            pF = p[o] * f + p[n] * (-f + 1)
            cL, cR = np.sqrt(g * T[o] * Rg), np.sqrt(g * T[n] * Rg)
            cF = cR * (1 - f) + cL * f
        nrm = N[sl]
        wp = A[sl] / VL[sl]
        UFN = UF[:, :, None] * nrm[:, None, :]
        Un = (UF * nrm).sum(axis=1, keepdims=True)
        np.add.at(gradU, o, UFN * wp[:, :, None]); np.add.at(divU, o, Un * wp)
        np.add.at(gradp, o, pF * nrm * wp); np.add.at(gradc, o, cF * nrm * wp)
        if neighbour:
            wn = -A[sl] / VR[sl]
            np.add.at(gradU, n, UFN * wn[:, :, None]); np.add.at(divU, n, Un * wn)
            np.add.at(gradp, n, pF * nrm * wn); np.add.at(gradc, n, cF * nrm * wn)

    faces(slice(0, Fi), True, False)
    for pid in P.sorted:                                           # postpro.py:538-547
        patch = P.patch[pid]
        s, n = patch["startFace"], patch["nFaces"]
        if n:
            faces(slice(s, s + n), False, patch["type"] not in O.COUPLED)
    if P.nRemoteCells > 0:
        faces(slice(P.nLocalFaces, P.nLocalFaces + P.nRemoteCells), False, False)
    return gradU, divU, gradp, gradc


def symmetrised_matrix(P, vtype, U, T, p, gradU, divU, gradp, gradc):
    """`getMaxEigenvalue` kernel, adFVM/postpro.py:551-655: MS = (M + M^T)/2 with M = M1/2 - M2 [C,5,5]."""
    C = P.nInternalCells
    g = P.gamma
    sg, g1 = np.sqrt(g), g - 1
    sg1 = np.sqrt(g1)
    U, T, p = U[:C], T[:C], p[:C]
    e = P.Cv * T                                                   # solver.conservative, density.py:173-181
    rho = p / (e * (g - 1))
    c = np.sqrt(g * p / rho)
    gradrho = g * (gradp - c * p) / (c * c)                        # postpro.py:561 (as written there)
    b, a = c / sg, sg1 * c / sg
    gradb, grada = gradc / sg, gradc * sg1 / sg
    Z = np.zeros((C, 1))
    col = lambda x, i: x[:, i:i + 1]
    gU = lambda i, j: gradU[:, i, j:j + 1]
    if vtype == "abarbanel":
        M1 = [divU, col(gradb, 0), col(gradb, 1), col(gradb, 2), Z,
              col(gradb, 0), divU, Z, Z, col(grada, 0),
              col(gradb, 1), Z, divU, Z, col(grada, 1),
              col(gradb, 2), Z, Z, divU, col(grada, 2),
              Z, col(grada, 0), col(grada, 1), col(grada, 2), divU]
        tmp1 = b * gradrho / rho
        tmp2 = a * gradp / (2 * p)
        tmp3 = 2 * grada / g1
        M2 = [Z, col(tmp1, 0), col(tmp1, 1), col(tmp1, 2), sg1 * divU / 2,
              Z, gU(0, 0), gU(0, 1), gU(0, 2), col(tmp2, 0),
              Z, gU(1, 0), gU(1, 1), gU(1, 2), col(tmp2, 1),
              Z, gU(2, 0), gU(2, 1), gU(2, 2), col(tmp2, 2),
              Z, col(tmp3, 0), col(tmp3, 1), col(tmp3, 2), g1 * divU / 2]
    elif vtype == "turkel":
        M1 = [divU, col(gradc, 0), col(gradc, 1), col(gradc, 2), Z,
              col(gradc, 0), divU, Z, Z, Z,
              col(gradc, 1), Z, divU, Z, Z,
              col(gradc, 2), Z, Z, divU, Z,
              Z, Z, Z, Z, divU]
        tmp1 = gradp / (rho * c)
        tmp2 = g1 * gradp / (2 * rho * c)
        tmp3 = gradp * PREF / (2 * g * p * rho * UREF)
        tmp4 = (gradp - c * c * gradrho) * UREF / PREF
        M2 = [g1 * divU / 2, col(tmp1, 0), col(tmp1, 1), col(tmp1, 2), divU * PREF / (2 * rho * c * UREF),
              col(tmp2, 0), gU(0, 0), gU(0, 1), gU(0, 2), col(tmp3, 0),
              col(tmp2, 1), gU(1, 0), gU(1, 1), gU(1, 2), col(tmp3, 1),
              col(tmp2, 2), gU(2, 0), gU(2, 1), gU(2, 2), col(tmp3, 2),
              Z, col(tmp4, 0), col(tmp4, 1), col(tmp4, 2), g1 * divU / 2]
    else:
        raise NotImplementedError("viscosity type %r" % (vtype,))
    M = (np.concatenate(M1, axis=1) / 2 - np.concatenate(M2, axis=1)).reshape(C, 5, 5)
    MS = (M + M.transpose(0, 2, 1)) / 2
    if vtype == "turkel":
        # Reference quirk, reproduced: adpy's kernel generator emits one store per DISTINCT output scalar
        # (adpy/adpy/tensor.py:358-359 `if op in names: continue`, output index looked up per op). In the turkel matrix
        # MS[0][0] and MS[4][4] are the same expression (divU/2 - (g-1) divU/2), so only [4][4] is stored and [0][0]
        # keeps the zero the output array was created with (24 stores in the generated kernel).
        MS[:, 0, 0] = 0.
    return MS


def adjoint_viscosity(spec, inputs, vtype, scaling, allreduce=None):
    """computeAdjointViscosity (adFVM/postpro.py:491-690) for the state in `inputs` (the positional inputs of `primal`).
    Returns M_2norm [nCells,1] (ghost rows filled like a default-boundary CellField) and DT [nFaces,1] (interp.central)."""
    import torch
    P = O.Problem(spec, inputs)
    C, nC = P.nInternalCells, P.nCells
    V = _np(P.volumes)
    if vtype == "uniform":                                         # postpro.py:660-661
        lam = np.ones((C, 1))
    else:
        U, T, p = _fields_with_ghosts(P)
        MS = symmetrised_matrix(P, vtype, U, T, p, *viscosity_gradients(P, U, T, p))
        lam = np.linalg.eigvalsh(MS)[:, 4:5]                       # scaling.cpp:84-106: dsyev, largest eigenvalue
    red = allreduce or (lambda x: x)
    Vt = red(V.sum())                                              # postpro.py:670-681
    Nrm = red((lam * lam * V / Vt).sum())
    M = lam * scaling / np.sqrt(Nrm)
    M = np.concatenate([M, np.zeros((nC - C, 1))], axis=0)
    M = _np(O.update_ghost(P, torch.as_tensor(M), "M_2norm", True))   # postpro.py:683-687: CellField with the default boundary
    own, nei, w = _np(P.owner), _np(P.neighbour), _np(P.weights)
    DT = M[own] * w + M[nei] * (-w + 1)                            # postpro.py:689-696
    return M, DT


def heat_matrix(P, DT, dt):
    """The system of `Matop::heat_equation` for ONE of the five fields (they share it: A0 = identity,
    matop_petsc.cpp:315-329), single rank: row i = (1 + sum_k a_ik) x_i - sum_k a_ik x_nbr(k) = u_i with
    a_ik = dt * DTF_f / V_i over the faces whose other cell is an internal cell (cellNeighboursMatOp > -1, cmesh.cpp:244-267),
    DTF = areas * DT / deltas (postpro.py:706-708)."""
    import scipy.sparse as sp
    C = P.nInternalCells
    V = _np(P.volumes)[:, 0]
    DTF = (_np(P.areas) * DT / _np(P.deltas))[:, 0]
    cf, cn = _np(P.cellFaces), _np(P.cellNeighbours)
    rows, cols, vals = [], [], []
    diag = np.ones(C)
    for k in range(6):
        nb, f = cn[:, k], cf[:, k]
        ok = nb < C                                               # ghost neighbours (any boundary face incl. cyclic) are not coupled
        a = DTF[f] * dt / V
        diag += np.where(ok, a, 0.)
        rows.append(np.nonzero(ok)[0]); cols.append(nb[ok]); vals.append(-a[ok])
    rows.append(np.arange(C)); cols.append(np.arange(C)); vals.append(diag)
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(C, C))


def apply_adjoint_viscosity(spec, inputs, DT, adj):
    """viscositySolver (adFVM/postpro.py:699-720): divide the adjoint fields by the cell volumes, one implicit diffusion step
    with face conductance DTF over the step's dt, multiply back. adj = [rhoa, rhoUa, rhoEa] (volume-weighted, as the
    reference's adjoint fields are); returns the three smoothed arrays."""
    from scipy.sparse.linalg import splu
    P = O.Problem(spec, inputs)
    V = _np(P.volumes)
    A = heat_matrix(P, DT, float(P.dt))
    u = np.concatenate([np.asarray(a, np.float64) for a in adj], axis=1) / V
    x = splu(A.tocsc()).solve(u)
    x = x * V
    return [x[:, 0:1].copy(), x[:, 1:4].copy(), x[:, 4:5].copy()]
