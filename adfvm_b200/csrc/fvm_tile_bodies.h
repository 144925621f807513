// Tiled flux kernels: one CTA per tile of T consecutive cells (fvm_tiles.h), one thread per face entry.
//
//   flux_tile       a8-a12: face flux (reconstruction + Riemann + viscous, adFVM/density.py:253-331) of every face
//                   touching the tile, scatter +F*A/V_owner / -F*A/V_neighbour (adFVM/op.py:12-29) into
//                   shared-memory accumulators, then RK stage update (adFVM/timestep.py:35-45) and the primitive
//                   conversion of the new state (adFVM/density.py:162-171) for the tile's cells.
//   flux_grad_tile  reverse of the flux + scatter: VJP of every face once, both sides' shares summed into
//                   shared-memory accumulators; ghost rows of boundary faces written directly (exclusive writer).
//
// Data movement of one CTA (R = scalar, W = threads = face entries per pass):
//   * face metrics live in HBM as per-pass CHUNKS [16][W] R + [W] u32 (area, n, 1/delta, dUnit, linW, quadW and the
//     packed entry word), laid out in the order the tile consumes them; one cp.async.bulk (TMA engine, SASS UBLKCP)
//     per pass streams a chunk into shared memory, completion on an mbarrier, issued one pass ahead of its use;
//   * qg [20][TS]: U(3),T,p and the 15 gradient components of the tile's own cells (slots [0,T): 20 bulk row copies of
//     the SoA arrays) and of its halo (slots [T,T+nHalo): cells of other tiles / ghost cells, gathered once);
//   * forward:  ivol [T] 1/V, acc [6][T] residual(5) + dtc          reverse:  r [5][TS] = abar*coef/V, acc [20][T].
// The per-face code therefore touches shared memory only, with compile-time strides and no branches on residency.
//
// Scatter order: entries are sorted by colour, faces of one colour never share an in-tile cell, and colours are
// applied one after the other (barrier in between) -> every cell receives its contributions in entry order.
// The CPU simulator (tests/hostsim) runs the same face/scatter/finish functions over the entries sequentially,
// which is the same order.
#pragma once
#include "fvm_bodies.h"
#if !defined(__CUDACC__)
#include <vector>
#endif

namespace fvm {

// scalars per chunk: 16 metric rows of W + W packed u32 words
template <typename R, int W> struct Chunk {
    static constexpr int kScalars = 16 * W + (W * 4) / (int)sizeof(R);
    static constexpr int kBytes = kScalars * (int)sizeof(R);
    FVM_HD static const unsigned* words(const R* chunk) { return reinterpret_cast<const unsigned*>(chunk + 16 * W); }
    FVM_HD static unsigned* words(R* chunk) { return reinterpret_cast<unsigned*>(chunk + 16 * W); }
};

struct TileEntry { int lo, ln, col, kind; bool valid; };
FVM_HD void tile_decode(unsigned w, TileEntry& e) {
    e.lo = (int)(w & 0x3FFu); e.ln = (int)((w >> 10) & 0x3FFu); e.col = (int)((w >> 20) & 0x1Fu);
    e.kind = (int)((w >> 25) & 3u); e.valid = ((w >> 27) & 1u) != 0;
}
template <typename R, int W> FVM_HD void tile_load_geom(const R* chunk, int i, Geom<R>& g) {
    const R* p = chunk + i;
    g.area = p[0]; g.n[0] = p[W]; g.n[1] = p[2 * W]; g.n[2] = p[3 * W]; g.idelta = p[4 * W];
    g.d[0] = p[5 * W]; g.d[1] = p[6 * W]; g.d[2] = p[7 * W]; g.lw[0] = p[8 * W]; g.lw[1] = p[9 * W];
    for (int k = 0; k < 3; k++) { g.qw[0][k] = p[(10 + k) * W]; g.qw[1][k] = p[(13 + k) * W]; }
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

// fills the chunks from the face-indexed metric arrays (mesh upload, one-off): slot i of the global slot list
template <typename R, int W> struct FillChunksBody {
    static constexpr const char* kName = "fill_chunks";
    const int* slot_face; const unsigned* slot_word; int sF;
    const R *area, *normal, *idelta, *dunit, *linw, *quadw;
    R* chunks;
    FVM_HD void operator()(int i) const {
        R* c = chunks + (long)(i / W) * Chunk<R, W>::kScalars;
        const int l = i % W, f = slot_face[i];
        Chunk<R, W>::words(c)[l] = slot_word[i];
        R v[16];
        if (f >= 0) {
            v[0] = area[f]; v[4] = idelta[f];
            for (int k = 0; k < 3; k++) { v[1 + k] = normal[(long)k * sF + f]; v[5 + k] = dunit[(long)k * sF + f]; }
            v[8] = linw[f]; v[9] = linw[(long)sF + f];
            for (int k = 0; k < 6; k++) v[10 + k] = quadw[(long)k * sF + f];
        } else {
            for (int k = 0; k < 16; k++) v[k] = R(0);
        }
        for (int k = 0; k < 16; k++) c[k * W + l] = v[k];
    }
};

#if defined(__CUDACC__)
// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// L2 prefetch of a row that a later phase of the CTA reads with ordinary loads (bytes: multiple of 16)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
    if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

// shared-memory carve-up common to both kernels: [qg 20*TS][extra ...][chunk][2 mbarriers]
template <typename R, int T, int TS, int W, int EXTRA> struct TileSmem {
    static constexpr size_t kChunkOff = ((size_t)(20 * TS + EXTRA) * sizeof(R) + 127) / 128 * 128;
    static constexpr size_t kBarOff = kChunkOff + Chunk<R, W>::kBytes;
    static constexpr size_t kBytes = kBarOff + 16;
};

// prologue: arm the barriers, start the bulk copies of the tile's own rows and of the first chunk, gather the halo rows
template <typename R, int T, int TS, int W>
__device__ __forceinline__ void tile_prologue(const MeshDev<R>& m, int t, const R* Q, const R* G, R* qg, R* chunk,
                                              unsigned long long* bars, int p0, int np) {
    const int c0 = t * T;
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        // rows may run past the last internal cell into ghost rows / padding (allocated, finite, unused)
        mbar_expect_tx(&bars[0], 20u * T * (unsigned)sizeof(R));
        for (int k = 0; k < 5; k++) bulk_g2s(qg + k * TS, Q + (long)k * m.sN + c0, T * (unsigned)sizeof(R), &bars[0]);
        for (int k = 0; k < 15; k++) bulk_g2s(qg + (5 + k) * TS, G + (long)k * m.sN + c0, T * (unsigned)sizeof(R), &bars[0]);
        if (np > 0) {
            mbar_expect_tx(&bars[1], (unsigned)Chunk<R, W>::kBytes);
            bulk_g2s(chunk, m.chunks + (long)p0 * Chunk<R, W>::kScalars, (unsigned)Chunk<R, W>::kBytes, &bars[1]);
        }
    }
    const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0;
    for (int h = tid; h < nh; h += nthr) tile_stage_cell<R, TS>(qg, T + h, Q, G, m.sN, m.halo_cell[h0 + h]);
}
#endif

// ------------------------------------------------------------------------------------------ forward
template <typename R, int T, int TS, int W> struct FluxTileBody {
    static constexpr const char* kName = "flux_tile";
    static constexpr int kThreads = W;
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G;            // this stage's primitives / gradients (ghosts filled)
    const R *W0, *W1, *W2;     // previous stage states (W1/W2 NULL when their alpha is 0)
    R a0, a1, a2, beta, dt;
    const R* S;                // source terms [5][sC]
    R* Wn;                     // new state
    R* Qn;                     // primitives of the new state (may be NULL)
    R* dtc_partial;            // [nTiles] per-tile max of dtc (may be NULL)
#if defined(__CUDACC__)
    typedef TileSmem<R, T, TS, W, 7 * T> Smem;
    static size_t smem_bytes() { return Smem::kBytes; }
#endif

    // flux per unit area of one entry + the scatter weights A/V of its in-tile sides
    FVM_HD void face(const Geom<R>& gm, const TileEntry& e, const R* qg, const R* ivol, Flux5<R>& F, R& wave, R& sO, R& sNb) const {
        Prim<R> qL, qR; Grad<R> gL, gR;
        tile_load_cell<R, TS>(qg, e.lo, qL, gL); tile_load_cell<R, TS>(qg, e.ln, qR, gR);
        face_flux(ph, e.kind, gm, qL, gL, qR, gR, F, wave);
        sO = e.lo < T ? gm.area * ivol[e.lo] : R(0);
        sNb = e.ln < T ? gm.area * ivol[e.ln] : R(0);
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
        for (int p = m.pass_start[t]; p < m.pass_start[t + 1]; p++) {
            const R* chunk = m.chunks + (long)p * Chunk<R, W>::kScalars;
            for (int i = 0; i < W; i++) {
                TileEntry e; tile_decode(Chunk<R, W>::words(chunk)[i], e);
                if (!e.valid) continue;
                Geom<R> gm; tile_load_geom<R, W>(chunk, i, gm);
                Flux5<R> F; R wave, sO, sNb;
                face(gm, e, qg.data(), ivol.data(), F, wave, sO, sNb);
                scatter(acc.data(), e.lo, e.ln, F, wave, sO, sNb);
            }
        }
        R mx = R(-1e30);
        for (int l = 0; l < nc; l++) { R d = finish(c0 + l, &acc[l]); mx = d > mx ? d : mx; }
        if (dtc_partial) dtc_partial[t] = mx;
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x;
        R* qg = reinterpret_cast<R*>(smem);
        R* ivol = qg + 20 * TS;
        R* acc = ivol + T;
        R* chunk = reinterpret_cast<R*>(smem + Smem::kChunkOff);
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + Smem::kBarOff);
        const int p0 = m.pass_start[t], np = m.pass_start[t + 1] - p0;
        tile_prologue<R, T, TS, W>(m, t, Q, G, qg, chunk, bars, p0, np);
        if (tid == 32) {      // the epilogue's inputs: bring them into L2 while the faces are computed
            const unsigned bytes = (unsigned)(nc * sizeof(R)) & ~15u;
            for (int k = 0; k < 5; k++) {
                bulk_prefetch_l2(W0 + (long)k * m.sC + c0, bytes); bulk_prefetch_l2(S + (long)k * m.sC + c0, bytes);
                if (W1) bulk_prefetch_l2(W1 + (long)k * m.sC + c0, bytes);
                if (W2) bulk_prefetch_l2(W2 + (long)k * m.sC + c0, bytes);
            }
        }
        for (int l = tid; l < T; l += W) ivol[l] = l < nc ? rcp(m.vol[c0 + l]) : R(0);
        for (int i = tid; i < 6 * T; i += W) acc[i] = R(0);
        mbar_wait(&bars[0], 0);
        __syncthreads();
        for (int p = 0; p < np; p++) {
            mbar_wait(&bars[1], (unsigned)(p & 1));
            TileEntry e; tile_decode(Chunk<R, W>::words(chunk)[tid], e);
            Geom<R> gm; tile_load_geom<R, W>(chunk, tid, gm);
            const int cfirst = (int)((Chunk<R, W>::words(chunk)[0] >> 20) & 0x1Fu), clast = (int)((Chunk<R, W>::words(chunk)[W - 1] >> 20) & 0x1Fu);
            __syncthreads();                       // every thread holds its metrics in registers: the chunk buffer is free
            if (tid == 0 && p + 1 < np) {
                fence_proxy_async();
                mbar_expect_tx(&bars[1], (unsigned)Chunk<R, W>::kBytes);
                bulk_g2s(chunk, m.chunks + (long)(p0 + p + 1) * Chunk<R, W>::kScalars, (unsigned)Chunk<R, W>::kBytes, &bars[1]);
            }
            Flux5<R> F; R wave = R(0), sO = R(0), sNb = R(0);
            if (e.valid) face(gm, e, qg, ivol, F, wave, sO, sNb);
            const int col = e.valid ? e.col : -1;
            for (int c = cfirst; c <= clast; c++) {
                if (col == c) scatter(acc, e.lo, e.ln, F, wave, sO, sNb);
                __syncthreads();
            }
        }
        R mx = R(-1e30);
        for (int l = tid; l < nc; l += W) { R d = finish(c0 + l, acc + l); mx = d > mx ? d : mx; }
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
template <typename R, int T, int TS, int W> struct FluxGradTileBody {
    static constexpr const char* kName = "flux_grad_tile";
    static constexpr int kThreads = W;
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G; const R* abar; R coef;
    R *Qb, *Gb;
#if defined(__CUDACC__)
    typedef TileSmem<R, T, TS, W, 5 * TS + 20 * T> Smem;
    static size_t smem_bytes() { return Smem::kBytes; }
#endif

    // r = abar*coef/V of an internal cell, 0 for ghost cells (boundary faces scatter to their owner only)
    FVM_HD void stage_r(R* r, int slot, int cell) const {
        if (cell < m.nInternalCells) { const R iv = coef * rcp(m.vol[cell]); for (int k = 0; k < 5; k++) r[k * TS + slot] = abar[(long)k * m.sC + cell] * iv; }
        else for (int k = 0; k < 5; k++) r[k * TS + slot] = R(0);
    }
    FVM_HD void face(const Geom<R>& gm, const TileEntry& e, const R* qg, const R* r, Prim<R>& qLb, Grad<R>& gLb, Prim<R>& qRb, Grad<R>& gRb) const {
        Prim<R> qL, qR; Grad<R> gL, gR;
        tile_load_cell<R, TS>(qg, e.lo, qL, gL); tile_load_cell<R, TS>(qg, e.ln, qR, gR);
        R d[5];
        for (int k = 0; k < 5; k++) d[k] = gm.area * (r[k * TS + e.lo] - r[k * TS + e.ln]);
        Flux5<R> Fb; Fb.rho = d[0]; Fb.rhoU[0] = d[1]; Fb.rhoU[1] = d[2]; Fb.rhoU[2] = d[3]; Fb.rhoE = d[4];
        zero(qLb); zero(gLb); zero(qRb); zero(gRb);
        face_flux_vjp(ph, e.kind, gm, qL, gL, qR, gR, Fb, qLb, gLb, qRb, gRb);
    }
    FVM_HD static void add20(R* a, const Prim<R>& q, const Grad<R>& g) {
        a[0] += q.U[0]; a[T] += q.U[1]; a[2 * T] += q.U[2]; a[3 * T] += q.T; a[4 * T] += q.p;
        for (int k = 0; k < 9; k++) a[(5 + k) * T] += g.U[k];
        for (int k = 0; k < 3; k++) { a[(14 + k) * T] += g.T[k]; a[(17 + k) * T] += g.p[k]; }
    }
    // ghost: global row of the ghost cell of a boundary face (its halo slot's cell), -1 for internal faces
    FVM_HD void scatter(R* acc, const TileEntry& e, int ghost, const Prim<R>& qLb, const Grad<R>& gLb, const Prim<R>& qRb, const Grad<R>& gRb) const {
        if (e.lo < T) add20(acc + e.lo, qLb, gLb);
        if (ghost >= 0) { store_prim(Qb, m.sN, ghost, qRb); store_grad(Gb, m.sN, ghost, gRb); }
        else if (e.ln < T) add20(acc + e.ln, qRb, gRb);
    }
    FVM_HD int ghost_of(int t, const TileEntry& e) const {
        if (e.ln < T) return -1;
        const int cell = m.halo_cell[m.halo_start[t] + e.ln - T];
        return cell >= m.nInternalCells ? cell : -1;
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
        for (int p = m.pass_start[t]; p < m.pass_start[t + 1]; p++) {
            const R* chunk = m.chunks + (long)p * Chunk<R, W>::kScalars;
            for (int i = 0; i < W; i++) {
                TileEntry e; tile_decode(Chunk<R, W>::words(chunk)[i], e);
                if (!e.valid) continue;
                Geom<R> gm; tile_load_geom<R, W>(chunk, i, gm);
                Prim<R> qLb, qRb; Grad<R> gLb, gRb;
                face(gm, e, qg.data(), r.data(), qLb, gLb, qRb, gRb);
                scatter(acc.data(), e, ghost_of(t, e), qLb, gLb, qRb, gRb);
            }
        }
        for (int l = 0; l < nc; l++) finish(c0 + l, &acc[l]);
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x;
        R* qg = reinterpret_cast<R*>(smem);
        R* r = qg + 20 * TS;
        R* acc = r + 5 * TS;
        R* chunk = reinterpret_cast<R*>(smem + Smem::kChunkOff);
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + Smem::kBarOff);
        const int p0 = m.pass_start[t], np = m.pass_start[t + 1] - p0;
        tile_prologue<R, T, TS, W>(m, t, Q, G, qg, chunk, bars, p0, np);
        for (int l = tid; l < T; l += W) { if (l < nc) stage_r(r, l, c0 + l); else for (int k = 0; k < 5; k++) r[k * TS + l] = R(0); }
        const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0;
        for (int h = tid; h < nh; h += W) stage_r(r, T + h, m.halo_cell[h0 + h]);
        for (int i = tid; i < 20 * T; i += W) acc[i] = R(0);
        mbar_wait(&bars[0], 0);
        __syncthreads();
        for (int p = 0; p < np; p++) {
            mbar_wait(&bars[1], (unsigned)(p & 1));
            TileEntry e; tile_decode(Chunk<R, W>::words(chunk)[tid], e);
            Geom<R> gm; tile_load_geom<R, W>(chunk, tid, gm);
            const int cfirst = (int)((Chunk<R, W>::words(chunk)[0] >> 20) & 0x1Fu), clast = (int)((Chunk<R, W>::words(chunk)[W - 1] >> 20) & 0x1Fu);
            __syncthreads();
            if (tid == 0 && p + 1 < np) {
                fence_proxy_async();
                mbar_expect_tx(&bars[1], (unsigned)Chunk<R, W>::kBytes);
                bulk_g2s(chunk, m.chunks + (long)(p0 + p + 1) * Chunk<R, W>::kScalars, (unsigned)Chunk<R, W>::kBytes, &bars[1]);
            }
            Prim<R> qLb, qRb; Grad<R> gLb, gRb;
            int ghost = -1;
            if (e.valid) { face(gm, e, qg, r, qLb, gLb, qRb, gRb); if (e.kind != (int)FACE_COUPLED || e.ln >= T) ghost = ghost_of(t, e); }
            const int col = e.valid ? e.col : -1;
            for (int c = cfirst; c <= clast; c++) {
                if (col == c) scatter(acc, e, ghost, qLb, gLb, qRb, gRb);
                __syncthreads();
            }
        }
        for (int l = tid; l < nc; l += W) finish(c0 + l, acc + l);
    }
#endif
};

}  // namespace fvm
