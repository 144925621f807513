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
#include "fvm_tile_bodies.h"

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
// The same product as a tile kernel (one CTA per tile of the flux kernels' plan, one thread per cell): the search direction of the
// tile's own cells arrives by bulk copy, that of its halo cells (face neighbours in other tiles, processor ghost rows) by the
// asynchronous gather, and the six-neighbour gather of five fields reads shared memory through the precomputed neighbour slots -
// the staging of GradAdjTileBody. Writes q and the tile's five partial dot products p . q (fixed-order tree over the CTA) to
// part [nTiles][5]; ViscPartialBody sums them.
template <typename R, int T, int TS> struct ViscSpmvTileBody {
    static constexpr const char* kName = "visc_cg_spmv";
    static constexpr int kThreads = T;
    static constexpr int kMinBlocks = 4;
    static constexpr int kPrefetchDistance = 148 * kMinBlocks;
    MeshDev<R> m; const R *cf, *dg, *p; R *q, *part;
#if defined(__CUDACC__)
    static constexpr size_t kRedOff = ((size_t)5 * TS * sizeof(R) + 15) / 16 * 16;
    static constexpr size_t kBarOff = kRedOff + (size_t)5 * (T / 32) * sizeof(R);
    static size_t smem_bytes() { return kBarOff + 16; }
#endif
    FVM_HD void cell(int c, int lo, const R* ps, R* y) const {
        const R d = dg[c];
        for (int k = 0; k < 5; k++) y[k] = d * ps[k * TS + lo];
        for (int j = 0; j < 6; j++) {
            const R v = cf[(long)j * m.sC + c];
            if (v == R(0)) continue;                       // uncoupled face (local boundary): its slot is never read
            const int sl = m.nbrSlot[(long)j * m.sC + c] & 0x7fff;
            for (int k = 0; k < 5; k++) y[k] -= v * ps[k * TS + sl];
        }
    }
#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        std::vector<R> ps((size_t)5 * TS, std::numeric_limits<R>::quiet_NaN());
        for (int l = 0; l < nc; l++) for (int k = 0; k < 5; k++) ps[(size_t)k * TS + l] = p[(long)k * m.sN + c0 + l];
        for (int h = m.halo_start[t]; h < m.halo_start[t + 1]; h++)
            for (int k = 0; k < 5; k++) ps[(size_t)k * TS + T + h - m.halo_start[t]] = p[(long)k * m.sN + m.halo_cell[h]];
        // the CTA's tree: per warp a shuffle tree over 32 lanes, then the warps in order
        R tot[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < T / 32; w++) {
            R lane[32][5];
            for (int l = 0; l < 32; l++) {
                const int i = w * 32 + l;
                for (int k = 0; k < 5; k++) lane[l][k] = R(0);
                if (i >= nc) continue;
                R y[5]; cell(c0 + i, i, ps.data(), y);
                for (int k = 0; k < 5; k++) { q[(long)k * m.sC + c0 + i] = y[k]; lane[l][k] = y[k] * ps[(size_t)k * TS + i]; }
            }
            for (int o = 16; o > 0; o >>= 1) for (int l = 0; l < o; l++) for (int k = 0; k < 5; k++) lane[l][k] += lane[l + o][k];
            for (int k = 0; k < 5; k++) tot[k] += lane[0][k];
        }
        for (int k = 0; k < 5; k++) part[(long)t * 5 + k] = tot[k];
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
        R* ps = reinterpret_cast<R*>(smem);
        R* red = reinterpret_cast<R*>(smem + kRedOff);
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kBarOff);
        unsigned long long *bar_rows = &bars[0], *bar_halo = &bars[1];
        const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0;
        constexpr unsigned kRow = T * (unsigned)sizeof(R);
        if (tid == 0) { mbar_init(bar_rows, 1); mbar_init(bar_halo, T); mbar_fence_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar_rows, 5u * kRow);
            for (int k = 0; k < 5; k++) bulk_g2s(ps + k * TS, p + (long)k * m.sN + c0, kRow, bar_rows);
        }
        {
            constexpr int kIter = (TS - T + T - 1) / T;
            int hc[kIter];
            #pragma unroll
            for (int i = 0; i < kIter; i++) { const int h = tid + i * T; hc[i] = h < nh ? m.halo_cell[h0 + h] : -1; }
            #pragma unroll
            for (int i = 0; i < kIter; i++) {
                const int cellh = hc[i], slot = T + tid + i * T;
                if (cellh < 0) continue;
                for (int k = 0; k < 5; k++) cp_async_elem<sizeof(R)>(ps + k * TS + slot, p + (long)k * m.sN + cellh);
            }
        }
        cp_async_arrive(bar_halo);
        {   // L2 prefetch of the rows of the tile that runs in this CTA slot one wave later
            const int tn = t + kPrefetchDistance;
            if (tn < m.nTiles && tid < 5) bulk_prefetch_l2(p + (long)tid * m.sN + (long)tn * T, kRow);
        }
        // per-cell inputs that do not depend on the staged rows, loaded before the wait
        R v[6], d = R(0); int sl[6];
        if (tid < nc) {
            d = dg[c0 + tid];
            #pragma unroll
            for (int j = 0; j < 6; j++) { v[j] = cf[(long)j * m.sC + c0 + tid]; sl[j] = m.nbrSlot[(long)j * m.sC + c0 + tid] & 0x7fff; }
        }
        mbar_wait(bar_rows, 0);
        mbar_wait(bar_halo, 0);
        R dot[5] = {R(0), R(0), R(0), R(0), R(0)};
        if (tid < nc) {
            R y[5];
            #pragma unroll
            for (int k = 0; k < 5; k++) y[k] = d * ps[k * TS + tid];
            #pragma unroll
            for (int j = 0; j < 6; j++) {
                if (v[j] == R(0)) continue;
                #pragma unroll
                for (int k = 0; k < 5; k++) y[k] -= v[j] * ps[k * TS + sl[j]];
            }
            #pragma unroll
            for (int k = 0; k < 5; k++) { q[(long)k * m.sC + c0 + tid] = y[k]; dot[k] = y[k] * ps[k * TS + tid]; }
        }
        #pragma unroll
        for (int k = 0; k < 5; k++) {
            R x = dot[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) red[w * 5 + k] = x;
        }
        __syncthreads();
        if (tid < 5) { R x = R(0); for (int i = 0; i < T / 32; i++) x += red[i * 5 + tid]; part[(long)t * 5 + tid] = x; }
    }
#endif
};
template <typename R> struct ViscPartialBody {
    static constexpr const char* kName = "visc_cg_dot";
    const R* part;
    FVM_HD V5<R> operator()(int t) const { V5<R> o; for (int k = 0; k < 5; k++) o.v[k] = part[(long)t * 5 + k]; return o; }
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
// the same update as a tile kernel (one CTA per tile, one thread per cell, nothing staged): every SM streams its own cells with
// all loads of a cell in flight at once, the five dot products leave as per-tile partial sums (fixed-order tree) in part [nTiles][5]
template <typename R, int T> struct ViscUpdateTileBody {
    static constexpr const char* kName = "visc_cg_update";
    static constexpr int kThreads = T;
    static constexpr int kMinBlocks = 8;
    MeshDev<R> m; const R *dg, *p, *q, *rz, *pq; R *x, *r, *part;
#if defined(__CUDACC__)
    static size_t smem_bytes() { return (size_t)5 * (T / 32) * sizeof(R); }
#endif
    FVM_HD void cell(int c, R* o) const {
        const R id = R(1) / dg[c];
        for (int k = 0; k < 5; k++) {
            const R al = pq[k] != R(0) ? rz[k] / pq[k] : R(0);
            x[(long)k * m.sN + c] += al * p[(long)k * m.sN + c];
            const R rr = r[(long)k * m.sC + c] - al * q[(long)k * m.sC + c];
            r[(long)k * m.sC + c] = rr; o[k] = rr * rr * id;
        }
    }
#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        R tot[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < T / 32; w++) {
            R lane[32][5];
            for (int l = 0; l < 32; l++) {
                for (int k = 0; k < 5; k++) lane[l][k] = R(0);
                if (w * 32 + l < nc) cell(c0 + w * 32 + l, lane[l]);
            }
            for (int o = 16; o > 0; o >>= 1) for (int l = 0; l < o; l++) for (int k = 0; k < 5; k++) lane[l][k] += lane[l + o][k];
            for (int k = 0; k < 5; k++) tot[k] += lane[0][k];
        }
        for (int k = 0; k < 5; k++) part[(long)t * 5 + k] = tot[k];
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
        R* red = reinterpret_cast<R*>(smem);
        R o[5] = {R(0), R(0), R(0), R(0), R(0)};
        if (tid < nc) cell(c0 + tid, o);
        #pragma unroll
        for (int k = 0; k < 5; k++) {
            R v = o[k];
            for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
            if (lane == 0) red[w * 5 + k] = v;
        }
        __syncthreads();
        if (tid < 5) { R v = R(0); for (int i = 0; i < T / 32; i++) v += red[i * 5 + tid]; part[(long)t * 5 + tid] = v; }
    }
#endif
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
