// Tiled flux kernels: one CTA per tile of T consecutive cells (fvm_tiles.h), one thread per face entry.
//
//   flux_tile       a8-a12: face flux (reconstruction + Riemann + viscous, adFVM/density.py:253-331) of every face
//                   touching the tile, scatter +F*A/V_owner / -F*A/V_neighbour (adFVM/op.py:12-29) into
//                   shared-memory accumulators, then RK stage update (adFVM/timestep.py:35-45) and the primitive
//                   conversion of the new state (adFVM/density.py:162-171) for the tile's cells.
//   flux_grad_tile  reverse of the flux + scatter: VJP of every face once, both sides' shares summed into
//                   shared-memory accumulators; ghost rows of boundary faces written directly (exclusive writer).
//
// Scatter order: entries are sorted by colour, faces of one colour never share an in-tile cell, and colours are
// applied one after the other (barrier in between) -> every cell receives its contributions in entry order.
// The CPU simulator walks the entries sequentially, which is the same order.
#pragma once
#include "fvm_bodies.h"
#if !defined(__CUDACC__)
#include <vector>
#endif

namespace fvm {

enum { kTileNone = 0x3FF };

struct TileFace { int f, lo, ln, col, o, n; };

template <typename R> FVM_HD void tile_entry(const MeshDev<R>& m, int e, TileFace& tf) {
    tf.f = m.ent_face[e];
    const unsigned w = m.ent_loc[e];
    tf.lo = (int)(w & 0x3FFu); tf.ln = (int)((w >> 10) & 0x3FFu); tf.col = (int)(w >> 20);
    tf.o = m.owner[tf.f]; tf.n = m.neigh[tf.f];
}

// cell rows either from the tile's staged copy in shared memory ([20][T]: U,T,p then the 15 gradient components)
// or from the global SoA arrays (cells of other tiles, ghost cells)
template <typename R> struct CellSrc {
    const R* Q; const R* G; int sN;
    const R* sQG; int T;
    FVM_HD void load(int cell, int local, Prim<R>& q, Grad<R>& g) const {
        if (sQG && local != kTileNone) {
            const R* p = sQG + local;
            q.U[0] = p[0]; q.U[1] = p[T]; q.U[2] = p[2 * T]; q.T = p[3 * T]; q.p = p[4 * T];
            for (int k = 0; k < 9; k++) g.U[k] = p[(5 + k) * T];
            for (int k = 0; k < 3; k++) { g.T[k] = p[(14 + k) * T]; g.p[k] = p[(17 + k) * T]; }
        } else {
            load_prim(Q, sN, cell, q); load_grad(G, sN, cell, g);
        }
    }
};

#if defined(__CUDACC__)
template <typename R> __device__ __forceinline__ R block_max(R v, R* scratch /* >= 32 */) {
    for (int o = 16; o > 0; o >>= 1) { R x = __shfl_down_sync(0xffffffffu, v, o); v = x > v ? x : v; }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        v = l < nw ? scratch[l] : R(-1e30);
        for (int o = 16; o > 0; o >>= 1) { R x = __shfl_down_sync(0xffffffffu, v, o); v = x > v ? x : v; }
    }
    return v;   // valid in thread 0
}
#endif

// ------------------------------------------------------------------------------------------ forward
template <typename R> struct FluxTileBody {
    static constexpr const char* kName = "flux_tile";
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G;            // this stage's primitives / gradients (ghosts filled)
    const R *W0, *W1, *W2;     // previous stage states (W1/W2 NULL when their alpha is 0)
    R a0, a1, a2, beta, dt;
    const R* S;                // source terms [5][sC]
    R* Wn;                     // new state
    R* Qn;                     // primitives of the new state (may be NULL)
    R* dtc_partial;            // [nTiles] per-tile max of dtc (may be NULL)
    static size_t smem_bytes(int T) { return (size_t)(20 + 1 + 6) * T * sizeof(R); }

    // flux per unit area of one entry + the scatter weights A/V of its in-tile sides (0 when not in the tile)
    FVM_HD void face(const TileFace& tf, const CellSrc<R>& src, const R* sIvol, Flux5<R>& F, R& wave, R& sO, R& sNb) const {
        Geom<R> gm; load_geom(m, tf.f, gm);
        Prim<R> qL, qR; Grad<R> gL, gR;
        src.load(tf.o, tf.lo, qL, gL); src.load(tf.n, tf.ln, qR, gR);
        face_flux(ph, face_kind(m, tf.f), gm, qL, gL, qR, gR, F, wave);
        sO = sNb = R(0);
        if (tf.lo != kTileNone) sO = gm.area * (sIvol ? sIvol[tf.lo] : R(1) / m.vol[tf.o]);
        if (tf.ln != kTileNone) sNb = gm.area * (sIvol ? sIvol[tf.ln] : R(1) / m.vol[tf.n]);
    }
    // RK stage update + primitives of the new state for cell c; res[k*rs] = residual component k, res[5*rs] = dtc
    FVM_HD R finish(int c, const R* res, int rs) const {
        R wn[5];
        for (int k = 0; k < 5; k++) {
            R v = a0 * W0[k * m.sC + c];
            if (W1) v += a1 * W1[k * m.sC + c];
            if (W2) v += a2 * W2[k * m.sC + c];
            v += -beta * (res[k * rs] - S[k * m.sC + c]) * dt;
            wn[k] = v;
            Wn[k * m.sC + c] = v;
        }
        if (Qn) { Prim<R> q; primitive(ph, wn[0], wn + 1, wn[4], q); store_prim(Qn, m.sN, c, q); }
        return res[5 * rs];
    }

#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int T = m.T, c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        std::vector<R> acc((size_t)6 * T, R(0));
        CellSrc<R> src{Q, G, m.sN, nullptr, T};
        for (int e = m.tile_start[t]; e < m.tile_start[t + 1]; e++) {
            TileFace tf; tile_entry(m, e, tf);
            Flux5<R> F; R wave, sO, sNb;
            face(tf, src, nullptr, F, wave, sO, sNb);
            const R fl[5] = {F.rho, F.rhoU[0], F.rhoU[1], F.rhoU[2], F.rhoE};
            if (tf.lo != kTileNone) { for (int k = 0; k < 5; k++) acc[k * T + tf.lo] += fl[k] * sO; acc[5 * T + tf.lo] += wave * sO; }
            if (tf.ln != kTileNone) { for (int k = 0; k < 5; k++) acc[k * T + tf.ln] += fl[k] * (-sNb); acc[5 * T + tf.ln] += wave * sNb; }
        }
        R mx = R(-1e30);
        for (int l = 0; l < nc; l++) { R d = finish(c0 + l, &acc[l], T); mx = d > mx ? d : mx; }
        if (dtc_partial) dtc_partial[t] = mx;
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int T = m.T, c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x, nthr = blockDim.x;
        R* sQG = reinterpret_cast<R*>(smem);
        R* sIvol = sQG + 20 * T;
        R* sAcc = sIvol + T;
        for (int i = tid; i < 20 * T; i += nthr) {
            const int k = i / T, l = i - k * T;
            if (l < nc) sQG[i] = (k < 5) ? Q[(long)k * m.sN + c0 + l] : G[(long)(k - 5) * m.sN + c0 + l];
        }
        for (int l = tid; l < T; l += nthr) sIvol[l] = l < nc ? R(1) / m.vol[c0 + l] : R(0);
        for (int i = tid; i < 6 * T; i += nthr) sAcc[i] = R(0);
        __syncthreads();
        CellSrc<R> src{Q, G, m.sN, sQG, T};
        const int e0 = m.tile_start[t], e1 = m.tile_start[t + 1];
        for (int base = e0; base < e1; base += nthr) {
            const int e = base + tid;
            const bool valid = e < e1;
            TileFace tf; Flux5<R> F; R wave = R(0), sO = R(0), sNb = R(0);
            tf.col = -1; tf.lo = tf.ln = kTileNone;
            if (valid) { tile_entry(m, e, tf); face(tf, src, sIvol, F, wave, sO, sNb); }
            const int cfirst = (int)(m.ent_loc[base] >> 20), clast = (int)(m.ent_loc[min(base + nthr, e1) - 1] >> 20);
            for (int c = cfirst; c <= clast; c++) {
                if (tf.col == c) {
                    if (tf.lo != kTileNone) {
                        R* a = sAcc + tf.lo;
                        a[0] += F.rho * sO; a[T] += F.rhoU[0] * sO; a[2 * T] += F.rhoU[1] * sO; a[3 * T] += F.rhoU[2] * sO;
                        a[4 * T] += F.rhoE * sO; a[5 * T] += wave * sO;
                    }
                    if (tf.ln != kTileNone) {
                        R* a = sAcc + tf.ln; const R s = -sNb;
                        a[0] += F.rho * s; a[T] += F.rhoU[0] * s; a[2 * T] += F.rhoU[1] * s; a[3 * T] += F.rhoU[2] * s;
                        a[4 * T] += F.rhoE * s; a[5 * T] += wave * sNb;
                    }
                }
                __syncthreads();
            }
        }
        R mx = R(-1e30);
        for (int l = tid; l < nc; l += nthr) { R d = finish(c0 + l, sAcc + l, T); mx = d > mx ? d : mx; }
        if (dtc_partial) {
            mx = block_max(mx, sQG);
            if (tid == 0) dtc_partial[t] = mx;
        }
    }
#endif
};

// ------------------------------------------------------------------------------------------ reverse
//   abar = adjoint of the stage OUTPUT state [5][sC]; coef = -beta_ii*dt (d W_new / d residual)
//   outputs Qb [5][sN], Gb [15][sN]: rows of internal cells and of the ghost cells of LOCAL+REMOTE boundary faces
template <typename R> struct FluxGradTileBody {
    static constexpr const char* kName = "flux_grad_tile";
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G; const R* abar; R coef;
    R *Qb, *Gb;
    static size_t smem_bytes(int T) { return (size_t)(20 + 5 + 20) * T * sizeof(R); }

    FVM_HD void rvec(int cell, int local, const R* sR, R* r) const {
        if (sR && local != kTileNone) { for (int k = 0; k < 5; k++) r[k] = sR[k * m.T + local]; }
        else { const R iv = coef / m.vol[cell]; for (int k = 0; k < 5; k++) r[k] = abar[(long)k * m.sC + cell] * iv; }
    }
    FVM_HD void face(const TileFace& tf, const CellSrc<R>& src, const R* sR, Prim<R>& qLb, Grad<R>& gLb, Prim<R>& qRb, Grad<R>& gRb) const {
        Geom<R> gm; load_geom(m, tf.f, gm);
        Prim<R> qL, qR; Grad<R> gL, gR;
        src.load(tf.o, tf.lo, qL, gL); src.load(tf.n, tf.ln, qR, gR);
        R rO[5], d[5];
        rvec(tf.o, tf.lo, sR, rO);
        if (tf.f < m.nInternalFaces) { R rN[5]; rvec(tf.n, tf.ln, sR, rN); for (int k = 0; k < 5; k++) d[k] = gm.area * (rO[k] - rN[k]); }
        else for (int k = 0; k < 5; k++) d[k] = gm.area * rO[k];
        Flux5<R> Fb; Fb.rho = d[0]; Fb.rhoU[0] = d[1]; Fb.rhoU[1] = d[2]; Fb.rhoU[2] = d[3]; Fb.rhoE = d[4];
        zero(qLb); zero(gLb); zero(qRb); zero(gRb);
        face_flux_vjp(ph, face_kind(m, tf.f), gm, qL, gL, qR, gR, Fb, qLb, gLb, qRb, gRb);
    }
    FVM_HD static void add20(R* a, int T, const Prim<R>& q, const Grad<R>& g) {
        a[0] += q.U[0]; a[T] += q.U[1]; a[2 * T] += q.U[2]; a[3 * T] += q.T; a[4 * T] += q.p;
        for (int k = 0; k < 9; k++) a[(5 + k) * T] += g.U[k];
        for (int k = 0; k < 3; k++) { a[(14 + k) * T] += g.T[k]; a[(17 + k) * T] += g.p[k]; }
    }
    FVM_HD void scatter(const TileFace& tf, R* acc, const Prim<R>& qLb, const Grad<R>& gLb, const Prim<R>& qRb, const Grad<R>& gRb) const {
        if (tf.lo != kTileNone) add20(acc + tf.lo, m.T, qLb, gLb);
        if (tf.f >= m.nInternalFaces) { store_prim(Qb, m.sN, tf.n, qRb); store_grad(Gb, m.sN, tf.n, gRb); }
        else if (tf.ln != kTileNone) add20(acc + tf.ln, m.T, qRb, gRb);
    }
    FVM_HD void finish(int c, const R* a, int T) const {
        for (int k = 0; k < 5; k++) Qb[(long)k * m.sN + c] = a[k * T];
        for (int k = 0; k < 15; k++) Gb[(long)k * m.sN + c] = a[(5 + k) * T];
    }

#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int T = m.T, c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        std::vector<R> acc((size_t)20 * T, R(0));
        CellSrc<R> src{Q, G, m.sN, nullptr, T};
        for (int e = m.tile_start[t]; e < m.tile_start[t + 1]; e++) {
            TileFace tf; tile_entry(m, e, tf);
            Prim<R> qLb, qRb; Grad<R> gLb, gRb;
            face(tf, src, nullptr, qLb, gLb, qRb, gRb);
            scatter(tf, acc.data(), qLb, gLb, qRb, gRb);
        }
        for (int l = 0; l < nc; l++) finish(c0 + l, &acc[l], T);
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int T = m.T, c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x, nthr = blockDim.x;
        R* sQG = reinterpret_cast<R*>(smem);
        R* sR = sQG + 20 * T;
        R* sAcc = sR + 5 * T;
        for (int i = tid; i < 20 * T; i += nthr) {
            const int k = i / T, l = i - k * T;
            if (l < nc) sQG[i] = (k < 5) ? Q[(long)k * m.sN + c0 + l] : G[(long)(k - 5) * m.sN + c0 + l];
            sAcc[i] = R(0);
        }
        for (int l = tid; l < nc; l += nthr) {
            const R iv = coef / m.vol[c0 + l];
            for (int k = 0; k < 5; k++) sR[k * T + l] = abar[(long)k * m.sC + c0 + l] * iv;
        }
        __syncthreads();
        CellSrc<R> src{Q, G, m.sN, sQG, T};
        const int e0 = m.tile_start[t], e1 = m.tile_start[t + 1];
        for (int base = e0; base < e1; base += nthr) {
            const int e = base + tid;
            const bool valid = e < e1;
            TileFace tf; Prim<R> qLb, qRb; Grad<R> gLb, gRb;
            tf.col = -1; tf.lo = tf.ln = kTileNone; tf.f = 0;
            if (valid) { tile_entry(m, e, tf); face(tf, src, sR, qLb, gLb, qRb, gRb); }
            const int cfirst = (int)(m.ent_loc[base] >> 20), clast = (int)(m.ent_loc[min(base + nthr, e1) - 1] >> 20);
            for (int c = cfirst; c <= clast; c++) {
                if (tf.col == c) scatter(tf, sAcc, qLb, gLb, qRb, gRb);
                __syncthreads();
            }
        }
        for (int l = tid; l < nc; l += nthr) finish(c0 + l, sAcc + l, T);
    }
#endif
};

}  // namespace fvm
