// Tiled flux kernels: one CTA per tile of T consecutive cells (fvm_tiles.h), one thread per face entry.
//
//   flux_tile       a8-a12: face flux (reconstruction + Riemann + viscous, adFVM/density.py:253-331) of every face
//                   touching the tile, scatter +F*A/V_owner / -F*A/V_neighbour (adFVM/op.py:12-29) into
//                   shared-memory accumulators, then RK stage update (adFVM/timestep.py:35-45) and the primitive
//                   conversion of the new state (adFVM/density.py:162-171) for the tile's cells.
//   flux_grad_tile  reverse of the flux + scatter: VJP of every face once, both sides' shares summed into
//                   shared-memory accumulators; ghost rows of boundary faces written directly (exclusive writer).
//
// Shared memory of a CTA (R = scalar):
//   qg  [20][TS]  U(3),T,p and the 15 gradient components of the tile's own cells (slots [0,T), coalesced rows of the
//                 SoA arrays) and of its halo (slots [T, T+nHalo): cells of other tiles / ghost cells, gathered once)
//   forward:  ivol [T] 1/V, acc [6][T] residual(5) + dtc          reverse:  r [5][TS] = abar*coef/V, acc [20][T]
// so the per-face code reads both of its cells from shared memory with compile-time strides and no branches.
//
// Scatter order: entries are sorted by colour, faces of one colour never share an in-tile cell, and colours are
// applied one after the other (barrier in between) -> every cell receives its contributions in entry order.
// The CPU simulator (tests/hostsim) runs the same stage/face/scatter/finish functions over the entries sequentially,
// which is the same order.
#pragma once
#include "fvm_bodies.h"
#if !defined(__CUDACC__)
#include <vector>
#endif

namespace fvm {

template <typename R> FVM_HD void tile_entry(const MeshDev<R>& m, int e, int& f, int& lo, int& ln, int& col) {
    f = m.ent_face[e];
    const unsigned w = m.ent_loc[e];
    lo = (int)(w & 0x3FFu); ln = (int)((w >> 10) & 0x3FFu); col = (int)(w >> 20);
}

template <typename R, int TS> FVM_HD void tile_load_cell(const R* qg, int l, Prim<R>& q, Grad<R>& g) {
    const R* p = qg + l;
    q.U[0] = p[0]; q.U[1] = p[TS]; q.U[2] = p[2 * TS]; q.T = p[3 * TS]; q.p = p[4 * TS];
    for (int k = 0; k < 9; k++) g.U[k] = p[(5 + k) * TS];
    for (int k = 0; k < 3; k++) { g.T[k] = p[(14 + k) * TS]; g.p[k] = p[(17 + k) * TS]; }
}
template <typename R, int TS> FVM_HD void tile_stage_cell(R* qg, int slot, const R* Q, const R* G, int sN, int cell) {
    for (int k = 0; k < 5; k++) qg[k * TS + slot] = Q[(long)k * sN + cell];
    for (int k = 0; k < 15; k++) qg[(5 + k) * TS + slot] = G[(long)k * sN + cell];
}

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
// rows of the tile's own cells (coalesced) and of its halo (gathered) -> qg
template <typename R, int T, int TS>
__device__ __forceinline__ void tile_stage_device(const MeshDev<R>& m, int t, const R* Q, const R* G, R* qg) {
    const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int l = tid; l < nc; l += nthr) tile_stage_cell<R, TS>(qg, l, Q, G, m.sN, c0 + l);
    const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0;
    for (int h = tid; h < nh; h += nthr) tile_stage_cell<R, TS>(qg, T + h, Q, G, m.sN, m.halo_cell[h0 + h]);
}
#endif

// ------------------------------------------------------------------------------------------ forward
template <typename R, int T, int TS> struct FluxTileBody {
    static constexpr const char* kName = "flux_tile";
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G;            // this stage's primitives / gradients (ghosts filled)
    const R *W0, *W1, *W2;     // previous stage states (W1/W2 NULL when their alpha is 0)
    R a0, a1, a2, beta, dt;
    const R* S;                // source terms [5][sC]
    R* Wn;                     // new state
    R* Qn;                     // primitives of the new state (may be NULL)
    R* dtc_partial;            // [nTiles] per-tile max of dtc (may be NULL)
    static size_t smem_bytes() { return (size_t)(20 * TS + T + 6 * T) * sizeof(R); }

    // flux per unit area of one entry + the scatter weights A/V of its in-tile sides
    FVM_HD void face(int f, int lo, int ln, const R* qg, const R* ivol, Flux5<R>& F, R& wave, R& sO, R& sNb) const {
        Geom<R> gm; load_geom(m, f, gm);
        Prim<R> qL, qR; Grad<R> gL, gR;
        tile_load_cell<R, TS>(qg, lo, qL, gL); tile_load_cell<R, TS>(qg, ln, qR, gR);
        face_flux(ph, face_kind(m, f), gm, qL, gL, qR, gR, F, wave);
        sO = lo < T ? gm.area * ivol[lo] : R(0);
        sNb = ln < T ? gm.area * ivol[ln] : R(0);
    }
    FVM_HD static void scatter(R* acc, int lo, int ln, const Flux5<R>& F, R wave, R sO, R sNb) {
        if (lo < T) {
            R* a = acc + lo;
            a[0] += F.rho * sO; a[T] += F.rhoU[0] * sO; a[2 * T] += F.rhoU[1] * sO; a[3 * T] += F.rhoU[2] * sO;
            a[4 * T] += F.rhoE * sO; a[5 * T] += wave * sO;
        }
        if (ln < T) {
            R* a = acc + ln; const R s = -sNb;
            a[0] += F.rho * s; a[T] += F.rhoU[0] * s; a[2 * T] += F.rhoU[1] * s; a[3 * T] += F.rhoU[2] * s;
            a[4 * T] += F.rhoE * s; a[5 * T] += wave * sNb;
        }
    }
    // RK stage update + primitives of the new state for cell c; res[k*T] = residual component k, res[5*T] = dtc
    FVM_HD R finish(int c, const R* res) const {
        R wn[5];
        for (int k = 0; k < 5; k++) {
            R v = a0 * W0[(long)k * m.sC + c];
            if (W1) v += a1 * W1[(long)k * m.sC + c];
            if (W2) v += a2 * W2[(long)k * m.sC + c];
            v += -beta * (res[k * T] - S[(long)k * m.sC + c]) * dt;
            wn[k] = v;
            Wn[(long)k * m.sC + c] = v;
        }
        if (Qn) { Prim<R> q; primitive(ph, wn[0], wn + 1, wn[4], q); store_prim(Qn, m.sN, c, q); }
        return res[5 * T];
    }

#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        std::vector<R> qg((size_t)20 * TS, R(0)), ivol(T, R(0)), acc((size_t)6 * T, R(0));
        for (int l = 0; l < nc; l++) { tile_stage_cell<R, TS>(qg.data(), l, Q, G, m.sN, c0 + l); ivol[l] = R(1) / m.vol[c0 + l]; }
        for (int h = m.halo_start[t]; h < m.halo_start[t + 1]; h++) tile_stage_cell<R, TS>(qg.data(), T + h - m.halo_start[t], Q, G, m.sN, m.halo_cell[h]);
        for (int e = m.tile_start[t]; e < m.tile_start[t + 1]; e++) {
            int f, lo, ln, col; tile_entry(m, e, f, lo, ln, col);
            Flux5<R> F; R wave, sO, sNb;
            face(f, lo, ln, qg.data(), ivol.data(), F, wave, sO, sNb);
            scatter(acc.data(), lo, ln, F, wave, sO, sNb);
        }
        R mx = R(-1e30);
        for (int l = 0; l < nc; l++) { R d = finish(c0 + l, &acc[l]); mx = d > mx ? d : mx; }
        if (dtc_partial) dtc_partial[t] = mx;
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x, nthr = blockDim.x;
        R* qg = reinterpret_cast<R*>(smem);
        R* ivol = qg + 20 * TS;
        R* acc = ivol + T;
        tile_stage_device<R, T, TS>(m, t, Q, G, qg);
        for (int l = tid; l < T; l += nthr) ivol[l] = l < nc ? R(1) / m.vol[c0 + l] : R(0);
        for (int i = tid; i < 6 * T; i += nthr) acc[i] = R(0);
        __syncthreads();
        const int e0 = m.tile_start[t], e1 = m.tile_start[t + 1];
        for (int base = e0; base < e1; base += nthr) {
            const int e = base + tid;
            int f = 0, lo = T, ln = T, col = -1;
            Flux5<R> F; R wave = R(0), sO = R(0), sNb = R(0);
            if (e < e1) { tile_entry(m, e, f, lo, ln, col); face(f, lo, ln, qg, ivol, F, wave, sO, sNb); }
            const int cfirst = (int)(m.ent_loc[base] >> 20), clast = (int)(m.ent_loc[min(base + nthr, e1) - 1] >> 20);
            for (int c = cfirst; c <= clast; c++) {
                if (col == c) scatter(acc, lo, ln, F, wave, sO, sNb);
                __syncthreads();
            }
        }
        R mx = R(-1e30);
        for (int l = tid; l < nc; l += nthr) { R d = finish(c0 + l, acc + l); mx = d > mx ? d : mx; }
        if (dtc_partial) {
            mx = block_max(mx, qg);
            if (tid == 0) dtc_partial[t] = mx;
        }
    }
#endif
};

// ------------------------------------------------------------------------------------------ reverse
//   abar = adjoint of the stage OUTPUT state [5][sC]; coef = -beta_ii*dt (d W_new / d residual)
//   outputs Qb [5][sN], Gb [15][sN]: rows of internal cells and of the ghost cells of ALL boundary faces
template <typename R, int T, int TS> struct FluxGradTileBody {
    static constexpr const char* kName = "flux_grad_tile";
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G; const R* abar; R coef;
    R *Qb, *Gb;
    static size_t smem_bytes() { return (size_t)(20 * TS + 5 * TS + 20 * T) * sizeof(R); }

    // r = abar*coef/V of an internal cell, 0 for ghost cells (boundary faces scatter to their owner only)
    FVM_HD void stage_r(R* r, int slot, int cell) const {
        if (cell < m.nInternalCells) { const R iv = coef / m.vol[cell]; for (int k = 0; k < 5; k++) r[k * TS + slot] = abar[(long)k * m.sC + cell] * iv; }
        else for (int k = 0; k < 5; k++) r[k * TS + slot] = R(0);
    }
    FVM_HD void face(int f, int lo, int ln, const R* qg, const R* r, Prim<R>& qLb, Grad<R>& gLb, Prim<R>& qRb, Grad<R>& gRb) const {
        Geom<R> gm; load_geom(m, f, gm);
        Prim<R> qL, qR; Grad<R> gL, gR;
        tile_load_cell<R, TS>(qg, lo, qL, gL); tile_load_cell<R, TS>(qg, ln, qR, gR);
        R d[5];
        for (int k = 0; k < 5; k++) d[k] = gm.area * (r[k * TS + lo] - r[k * TS + ln]);
        Flux5<R> Fb; Fb.rho = d[0]; Fb.rhoU[0] = d[1]; Fb.rhoU[1] = d[2]; Fb.rhoU[2] = d[3]; Fb.rhoE = d[4];
        zero(qLb); zero(gLb); zero(qRb); zero(gRb);
        face_flux_vjp(ph, face_kind(m, f), gm, qL, gL, qR, gR, Fb, qLb, gLb, qRb, gRb);
    }
    FVM_HD static void add20(R* a, const Prim<R>& q, const Grad<R>& g) {
        a[0] += q.U[0]; a[T] += q.U[1]; a[2 * T] += q.U[2]; a[3 * T] += q.T; a[4 * T] += q.p;
        for (int k = 0; k < 9; k++) a[(5 + k) * T] += g.U[k];
        for (int k = 0; k < 3; k++) { a[(14 + k) * T] += g.T[k]; a[(17 + k) * T] += g.p[k]; }
    }
    FVM_HD void scatter(R* acc, int f, int lo, int ln, const Prim<R>& qLb, const Grad<R>& gLb, const Prim<R>& qRb, const Grad<R>& gRb) const {
        if (lo < T) add20(acc + lo, qLb, gLb);
        if (f >= m.nInternalFaces) { const int g = m.nInternalCells + (f - m.nInternalFaces); store_prim(Qb, m.sN, g, qRb); store_grad(Gb, m.sN, g, gRb); }
        else if (ln < T) add20(acc + ln, qRb, gRb);
    }
    FVM_HD void finish(int c, const R* a) const {
        for (int k = 0; k < 5; k++) Qb[(long)k * m.sN + c] = a[k * T];
        for (int k = 0; k < 15; k++) Gb[(long)k * m.sN + c] = a[(5 + k) * T];
    }

#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        std::vector<R> qg((size_t)20 * TS, R(0)), r((size_t)5 * TS, R(0)), acc((size_t)20 * T, R(0));
        for (int l = 0; l < nc; l++) { tile_stage_cell<R, TS>(qg.data(), l, Q, G, m.sN, c0 + l); stage_r(r.data(), l, c0 + l); }
        for (int h = m.halo_start[t]; h < m.halo_start[t + 1]; h++) {
            const int slot = T + h - m.halo_start[t];
            tile_stage_cell<R, TS>(qg.data(), slot, Q, G, m.sN, m.halo_cell[h]); stage_r(r.data(), slot, m.halo_cell[h]);
        }
        for (int e = m.tile_start[t]; e < m.tile_start[t + 1]; e++) {
            int f, lo, ln, col; tile_entry(m, e, f, lo, ln, col);
            Prim<R> qLb, qRb; Grad<R> gLb, gRb;
            face(f, lo, ln, qg.data(), r.data(), qLb, gLb, qRb, gRb);
            scatter(acc.data(), f, lo, ln, qLb, gLb, qRb, gRb);
        }
        for (int l = 0; l < nc; l++) finish(c0 + l, &acc[l]);
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x, nthr = blockDim.x;
        R* qg = reinterpret_cast<R*>(smem);
        R* r = qg + 20 * TS;
        R* acc = r + 5 * TS;
        tile_stage_device<R, T, TS>(m, t, Q, G, qg);
        for (int l = tid; l < nc; l += nthr) stage_r(r, l, c0 + l);
        const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0;
        for (int h = tid; h < nh; h += nthr) stage_r(r, T + h, m.halo_cell[h0 + h]);
        for (int i = tid; i < 20 * T; i += nthr) acc[i] = R(0);
        __syncthreads();
        const int e0 = m.tile_start[t], e1 = m.tile_start[t + 1];
        for (int base = e0; base < e1; base += nthr) {
            const int e = base + tid;
            int f = 0, lo = T, ln = T, col = -1;
            Prim<R> qLb, qRb; Grad<R> gLb, gRb;
            if (e < e1) { tile_entry(m, e, f, lo, ln, col); face(f, lo, ln, qg, r, qLb, gLb, qRb, gRb); }
            const int cfirst = (int)(m.ent_loc[base] >> 20), clast = (int)(m.ent_loc[min(base + nthr, e1) - 1] >> 20);
            for (int c = cfirst; c <= clast; c++) {
                if (col == c) scatter(acc, f, lo, ln, qLb, gLb, qRb, gRb);
                __syncthreads();
            }
        }
        for (int l = tid; l < nc; l += nthr) finish(c0 + l, acc + l);
    }
#endif
};

}  // namespace fvm
