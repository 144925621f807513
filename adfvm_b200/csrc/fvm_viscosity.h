// Adjoint artificial viscosity: the stabilisation `apps/adjoint.py:127-141,288-289` applies to the
// adjoint fields every `viscousInterval` steps when the case file sets adjParams = [scaling, type, None].
//
//   computeAdjointViscosity  adFVM/postpro.py:491-697   M_2norm = largest eigenvalue of the symmetrised 5x5 matrix
//                            adFVM/cpp/scaling.cpp:84-106  built from div U, grad U, grad p, grad c of the step's start state
//                                                        (LAPACK dsyev / cusolver syevjBatched in the reference), normalised
//                                                        by its volume-weighted RMS and multiplied by `scaling`
//   viscositySolver          adFVM/postpro.py:699-720   one backward-Euler diffusion step of the five adjoint fields
//                            adFVM/cpp/matop_petsc.cpp:288-372 (assembly), matop_cuda.cpp:160-231
//
// Here: one thread per cell for the matrix + a cyclic Jacobi eigenvalue iteration in registers (ViscEigBody; gradU, gradp are
// the Green-Gauss gradients G of the forward sweep - the reference's `gradients` kernel is the same sum, postpro.py:499-524 -
// grad c is gathered from the six neighbours), and a Jacobi-preconditioned conjugate-gradient solve of the five right-hand
// sides at once on the volume-weighted (symmetric) form of the system, matrix-free over the cell-neighbour lists, scalars
// kept on the device (ViscSpmvBody / ViscUpdateBody / ViscDirBody). The reference's CUDA build runs 1000 Jacobi sweeps per
// field instead (matop_cuda.cpp:205), its CPU build GMRES + hypre; all three approximate the same linear system.
#pragma once
#include "fvm_bodies.h"

namespace fvm {

enum ViscType { VISC_NONE = 0, VISC_ABARBANEL = 1, VISC_TURKEL = 2, VISC_UNIFORM = 3 };

template <typename R> struct V5 { R v[5]; };
template <typename R> FVM_HD V5<R> v5_zero() { V5<R> z; for (int k = 0; k < 5; k++) z.v[k] = R(0); return z; }
template <typename R> FVM_HD V5<R> v5_add(const V5<R>& a, const V5<R>& b) { V5<R> z; for (int k = 0; k < 5; k++) z.v[k] = a.v[k] + b.v[k]; return z; }

// largest eigenvalue of the symmetric matrix a (both triangles given; destroyed): cyclic Jacobi sweeps until the off-diagonal
// part is below rounding. Fully unrolled index loops keep the 25 entries in registers on the device.
template <typename R> FVM_HD R sym5_max_eig(R (&a)[5][5]) {
    const R eps = sizeof(R) == 8 ? R(1e-32) : R(1e-14);      // on the SQUARED off-diagonal norm, relative to the squared diagonal
    for (int sweep = 0; sweep < 30; sweep++) {
        R off = R(0), dg = R(0);
#if defined(__CUDA_ARCH__)
        #pragma unroll
#endif
        for (int p = 0; p < 5; p++) {
            dg += a[p][p] * a[p][p];
#if defined(__CUDA_ARCH__)
            #pragma unroll
#endif
            for (int q = 0; q < 5; q++) if (q > p) off += a[p][q] * a[p][q];
        }
        if (off <= eps * (dg + off)) break;
#if defined(__CUDA_ARCH__)
        #pragma unroll
#endif
        for (int p = 0; p < 4; p++) {
#if defined(__CUDA_ARCH__)
            #pragma unroll
#endif
            for (int q = 1; q < 5; q++) {
                if (q <= p) continue;
                const R apq = a[p][q];
                if (apq == R(0)) continue;
                const R theta = (a[q][q] - a[p][p]) / (R(2) * apq);
                const R t = (theta < R(0) ? R(-1) : R(1)) / (fabs(theta) + sqrt(theta * theta + R(1)));
                const R c = R(1) / sqrt(t * t + R(1)), s = t * c;
                a[p][p] -= t * apq; a[q][q] += t * apq; a[p][q] = a[q][p] = R(0);
#if defined(__CUDA_ARCH__)
                #pragma unroll
#endif
                for (int r = 0; r < 5; r++) {
                    if (r == p || r == q) continue;
                    const R arp = a[r][p], arq = a[r][q];
                    a[r][p] = a[p][r] = c * arp - s * arq;
                    a[r][q] = a[q][r] = s * arp + c * arq;
                }
            }
        }
    }
    R mx = a[0][0];
    for (int p = 1; p < 5; p++) mx = a[p][p] > mx ? a[p][p] : mx;
    return mx;
}

// M1/2 - M2 of `getMaxEigenvalue` (adFVM/postpro.py:551-655), symmetrised. gU[3*i+j] = dU_i/dx_j.
template <typename R> FVM_HD void visc_matrix(const Phys<R>& ph, int type, R T, R p, const R* gU, const R* gp, const R* gc, R (&MS)[5][5]) {
    const R g = ph.gamma, g1 = ph.gm1;
    const R divU = gU[0] + gU[4] + gU[8];
    const R rho = p / (ph.Cv * T * g1);                       // solver.conservative, adFVM/density.py:173-181
    const R c = sqrt(g * p / rho);
    R grho[3];
    for (int k = 0; k < 3; k++) grho[k] = g * (gp[k] - c * p) / (c * c);      // postpro.py:561, as written there
    R M[5][5];
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) M[i][j] = R(0);
    if (type == VISC_ABARBANEL) {
        const R sg = sqrt(g), sg1 = sqrt(g1);
        const R b = c / sg, a = sg1 * c / sg;
        for (int k = 0; k < 3; k++) {
            const R gb = gc[k] / sg, ga = gc[k] * sg1 / sg;
            // M1/2
            M[0][1 + k] += gb / 2; M[1 + k][0] += gb / 2; M[1 + k][4] += ga / 2; M[4][1 + k] += ga / 2;
            // - M2
            M[0][1 + k] -= b * grho[k] / rho;
            M[1 + k][4] -= a * gp[k] / (2 * p);
            M[4][1 + k] -= 2 * ga / g1;
            for (int j = 0; j < 3; j++) M[1 + k][1 + j] -= gU[3 * k + j];
        }
        for (int i = 0; i < 5; i++) M[i][i] += divU / 2;
        M[0][4] -= sg1 * divU / 2;
        M[4][4] -= g1 * divU / 2;
    } else {                                                   // VISC_TURKEL
        const R Uref = R(33.), pref = R(1e5);                  // adFVM/density.py:57-59
        for (int k = 0; k < 3; k++) {
            M[0][1 + k] += gc[k] / 2; M[1 + k][0] += gc[k] / 2;
            M[0][1 + k] -= gp[k] / (rho * c);
            M[1 + k][0] -= g1 * gp[k] / (2 * rho * c);
            M[1 + k][4] -= gp[k] * pref / (2 * g * p * rho * Uref);
            M[4][1 + k] -= (gp[k] - c * c * grho[k]) * Uref / pref;
            for (int j = 0; j < 3; j++) M[1 + k][1 + j] -= gU[3 * k + j];
        }
        for (int i = 0; i < 5; i++) M[i][i] += divU / 2;
        M[0][0] -= g1 * divU / 2;
        M[0][4] -= divU * pref / (2 * rho * c * Uref);
        M[4][4] -= g1 * divU / 2;
    }
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) MS[i][j] = (M[i][j] + M[j][i]) / 2;
    // Reference quirk, reproduced: adpy's kernel generator stores each DISTINCT output scalar once (adpy/adpy/tensor.py:358-359);
    // in the turkel matrix MS[0][0] and MS[4][4] are the same expression, only [4][4] is stored and [0][0] stays zero
    if (type == VISC_TURKEL) MS[0][0] = R(0);
}

// per cell: largest eigenvalue of the symmetrised matrix. Q, G: primitives (ghost rows filled) / Green-Gauss gradients of the state
template <typename R> struct ViscEigBody {
    static constexpr const char* kName = "visc_eig";
    Phys<R> ph; MeshDev<R> m; int type; const R* Q; const R* G; R* lam;
    FVM_HD void operator()(int c) const {
        if (type == VISC_UNIFORM) { lam[c] = R(1); return; }  // postpro.py:660-661
        const int sN = m.sN, sC = m.sC;
        const R T = Q[3 * sN + c], p = Q[4 * sN + c];
        R gU[9], gp[3], gc[3] = {R(0), R(0), R(0)};
        for (int k = 0; k < 9; k++) gU[k] = G[k * sN + c];
        for (int k = 0; k < 3; k++) gp[k] = G[(12 + k) * sN + c];
        const R gR = ph.gamma * (ph.Cp - ph.Cv);
        const R cc = sqrt(gR * T);
        for (int j = 0; j < 6; j++) {                          // grad c: central face values of c = sqrt(gamma R T), postpro.py:512-515
            const int nb = m.cellNbr[j * sC + c];
            const R a = m.cfm[(long)(4 * j + 3) * sC + c];
            const R cf = cc * a + sqrt(gR * Q[3 * sN + nb]) * (R(1) - a);
            for (int k = 0; k < 3; k++) gc[k] += cf * m.cfm[(long)(4 * j + k) * sC + c];
        }
        const R iv = R(1) / m.vol[c];
        for (int k = 0; k < 3; k++) gc[k] *= iv;
        R MS[5][5];
        visc_matrix(ph, type, T, p, gU, gp, gc, MS);
        lam[c] = sym5_max_eig(MS);
    }
};
template <typename R> struct ViscVolBody { static constexpr const char* kName = "visc_vol"; const R* vol; FVM_HD R operator()(int c) const { return vol[c]; } };
template <typename R> struct ViscNormBody { static constexpr const char* kName = "visc_norm"; const R* vol; const R* lam; FVM_HD R operator()(int c) const { return lam[c] * lam[c] * vol[c]; } };
// M_2norm = lam * scaling / sqrt(sum lam^2 V / sum V) (postpro.py:670-681); s[0] = sum V, s[1] = sum lam^2 V (over all ranks)
template <typename R> struct ViscScaleBody {
    static constexpr const char* kName = "visc_scale";
    const R* lam; const R* s; R scaling; R* M;
    FVM_HD void operator()(int c) const { M[c] = lam[c] * scaling / sqrt(s[1] / s[0]); }
};
// ghost rows of a scalar cell field with the mesh's default boundary (cyclic copy, otherwise zeroGradient), local patches
template <typename R> struct GhostScalarBody {
    static constexpr const char* kName = "ghost_scalar";
    MeshDev<R> m; R* X;
    FVM_HD void operator()(int b) const {
        const int f = m.nInternalFaces + b;
        const PatchDev<R>& P = m.patches[m.bpatch[b]];
        const int s = (P.gbc == BC_CYCLIC) ? m.owner[P.nbrStartFace + (f - P.startFace)] : m.owner[f];
        X[m.nInternalCells + b] = X[s];
    }
};
// DT = interp.central(M_2norm) on every face (postpro.py:689-696); reference face order is restored by the caller
template <typename R> struct ViscFaceBody {
    static constexpr const char* kName = "visc_face";
    MeshDev<R> m; const R* M; R* DT;
    FVM_HD void operator()(int f) const { const R w = m.weight[f]; DT[f] = M[m.owner[f]] * w + M[m.neigh[f]] * (R(1) - w); }
};
// volume-weighted system of Matop::heat_equation (matop_petsc.cpp:340-372): V_i x_i + sum_k cf_ik (x_i - x_nbr) = V_i u_i with
// cf_ik = dt * areas * DT / deltas of the face (symmetric), over the faces whose other cell is an internal cell of this or -
// through a processor patch - of another rank. cf [6][sC], dg [sC] = V_i + sum_k cf_ik.
template <typename R> struct ViscCoefBody {
    static constexpr const char* kName = "visc_coef";
    MeshDev<R> m; const R* M; R dt; R* cf; R* dg;
    FVM_HD void operator()(int c) const {
        R d = m.vol[c];
        for (int j = 0; j < 6; j++) {
            const int f = m.cellFaces[j * m.sC + c], nb = m.cellNbr[j * m.sC + c];
            const bool coupled = nb < m.nInternalCells || nb >= m.nLocalCells;
            const R w = m.weight[f];
            const R DT = M[m.owner[f]] * w + M[m.neigh[f]] * (R(1) - w);
            const R v = coupled ? dt * m.area[f] * DT * m.idelta[f] : R(0);
            cf[(long)j * m.sC + c] = v; d += v;
        }
        dg[c] = d;
    }
};
template <typename R> FVM_HD void visc_apply(const MeshDev<R>& m, const R* cf, const R* dg, const R* x, int sX, int c, R* y) {
    for (int k = 0; k < 5; k++) y[k] = dg[c] * x[(long)k * sX + c];
    for (int j = 0; j < 6; j++) {
        const R v = cf[(long)j * m.sC + c];
        if (v == R(0)) continue;
        const int nb = m.cellNbr[j * m.sC + c];
        for (int k = 0; k < 5; k++) y[k] -= v * x[(long)k * sX + nb];
    }
}
// x0 = adj / V (the reference starts its solvers from the right-hand side u = adj/V too)
template <typename R> struct ViscStartBody {
    static constexpr const char* kName = "visc_start";
    MeshDev<R> m; const R* adj; R* x;
    FVM_HD void operator()(int c) const { const R iv = R(1) / m.vol[c]; for (int k = 0; k < 5; k++) x[(long)k * m.sN + c] = adj[(long)k * m.sC + c] * iv; }
};
// r = adj - A x0, p = r / dg; returns (r . r/dg) per field; bz (second reduction): adj . adj/dg, the scale of the stopping test
template <typename R> struct ViscInitBody {
    static constexpr const char* kName = "visc_cg_init";
    MeshDev<R> m; const R *cf, *dg, *adj, *x; R *r, *p;
    FVM_HD V5<R> operator()(int c) const {
        R y[5]; visc_apply(m, cf, dg, x, m.sN, c, y);
        V5<R> o; const R id = R(1) / dg[c];
        for (int k = 0; k < 5; k++) { const R rr = adj[(long)k * m.sC + c] - y[k]; r[(long)k * m.sC + c] = rr; p[(long)k * m.sN + c] = rr * id; o.v[k] = rr * rr * id; }
        return o;
    }
};
template <typename R> struct ViscRhsNormBody {
    static constexpr const char* kName = "visc_cg_rhs";
    MeshDev<R> m; const R *dg, *adj;
    FVM_HD V5<R> operator()(int c) const { V5<R> o; const R id = R(1) / dg[c]; for (int k = 0; k < 5; k++) { const R a = adj[(long)k * m.sC + c]; o.v[k] = a * a * id; } return o; }
};
// q = A p; returns p . q
template <typename R> struct ViscSpmvBody {
    static constexpr const char* kName = "visc_cg_spmv";
    MeshDev<R> m; const R *cf, *dg, *p; R* q;
    FVM_HD V5<R> operator()(int c) const {
        R y[5]; visc_apply(m, cf, dg, p, m.sN, c, y);
        V5<R> o;
        for (int k = 0; k < 5; k++) { q[(long)k * m.sC + c] = y[k]; o.v[k] = y[k] * p[(long)k * m.sN + c]; }
        return o;
    }
};
// alpha = rz / pq (device scalars); x += alpha p, r -= alpha q; returns r . r/dg
template <typename R> struct ViscUpdateBody {
    static constexpr const char* kName = "visc_cg_update";
    MeshDev<R> m; const R *dg, *p, *q, *rz, *pq; R *x, *r;
    FVM_HD V5<R> operator()(int c) const {
        V5<R> o; const R id = R(1) / dg[c];
        for (int k = 0; k < 5; k++) {
            const R al = pq[k] != R(0) ? rz[k] / pq[k] : R(0);
            x[(long)k * m.sN + c] += al * p[(long)k * m.sN + c];
            const R rr = r[(long)k * m.sC + c] - al * q[(long)k * m.sC + c];
            r[(long)k * m.sC + c] = rr; o.v[k] = rr * rr * id;
        }
        return o;
    }
};
// beta = rz_new / rz; p = r/dg + beta p
template <typename R> struct ViscDirBody {
    static constexpr const char* kName = "visc_cg_dir";
    MeshDev<R> m; const R *dg, *r, *rz, *rzn; R* p;
    FVM_HD void operator()(int c) const {
        const R id = R(1) / dg[c];
        for (int k = 0; k < 5; k++) {
            const R be = rz[k] != R(0) ? rzn[k] / rz[k] : R(0);
            p[(long)k * m.sN + c] = r[(long)k * m.sC + c] * id + be * p[(long)k * m.sN + c];
        }
    }
};
// adjoint fields back in their volume-weighted form (multiplyFields, postpro.py:714-719)
template <typename R> struct ViscFinishBody {
    static constexpr const char* kName = "visc_finish";
    MeshDev<R> m; const R* x; R* adj;
    FVM_HD void operator()(int c) const { const R v = m.vol[c]; for (int k = 0; k < 5; k++) adj[(long)k * m.sC + c] = x[(long)k * m.sN + c] * v; }
};

}  // namespace fvm
