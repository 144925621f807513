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

// Ampere-style asynchronous element copies (SASS LDGSTS) for the halo gather: no registers, no stall at issue
template <int BYTES> __device__ __forceinline__ void cp_async_elem(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory carve-up common to both kernels: [qg 20*TS][extra ...][chunk][3 mbarriers]
template <typename R, int T, int TS, int W, int EXTRA> struct TileSmem {
    static constexpr size_t kChunkOff = ((size_t)(20 * TS + EXTRA) * sizeof(R) + 127) / 128 * 128;
    static constexpr size_t kBarOff = kChunkOff + Chunk<R, W>::kBytes;
    static constexpr size_t kBytes = kBarOff + 32;
};
#endif

// ------------------------------------------------------------------------------------------ forward
template <typename R, int T, int TS, int W> struct FluxTileBody {
    static constexpr const char* kName = "flux_tile";
    static constexpr int kThreads = W;
    static constexpr int kMinBlocks = (sizeof(R) == 8 && W == 160) ? 3 : 1;
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G;            // this stage's primitives / gradients (ghosts filled)
    const R *W0, *W1, *W2;     // previous stage states (W1/W2 NULL when their alpha is 0; never both set in SSPRK3)
    R a0, a1, a2, beta, dt;
    const R* S;                // source terms [5][sC]
    R* Wn;                     // new state
    R* Qn;                     // primitives of the new state (may be NULL)
    R* dtc_partial;            // [nTiles] per-tile max of dtc (may be NULL)
#if defined(__CUDACC__)
    typedef TileSmem<R, T, TS, W, 7 * T> Smem;
    static size_t smem_bytes() { return Smem::kBytes; }
#endif

    // flux per unit area of one entry + the scatter weights A/V of its in-tile sides (vol: the tile's own volumes [T])
    FVM_HD void face(const Geom<R>& gm, const TileEntry& e, const R* qg, const R* vol, Flux5<R>& F, R& wave, R& sO, R& sNb) const {
        Prim<R> qL, qR; Grad<R> gL, gR;
        tile_load_cell<R, TS>(qg, e.lo, qL, gL); tile_load_cell<R, TS>(qg, e.ln, qR, gR);
        face_flux(ph, e.kind, gm, qL, gL, qR, gR, F, wave);
        sO = e.lo < T ? gm.area * rcp(vol[e.lo]) : R(0);
        sNb = e.ln < T ? gm.area * rcp(vol[e.ln]) : R(0);
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
    // RK stage update + primitives of the new state for cell c; res[k*T] = residual component k, res[5*T] = dtc;
    // w0/wx/src: this cell's previous-stage states and source, component stride ws (wx = W1 or W2, coefficient ax)
    FVM_HD R finish(int c, const R* res, const R* w0, const R* wx, R ax, const R* src, int ws) const {
        R wn[5];
        for (int k = 0; k < 5; k++) {
            R v = a0 * w0[(long)k * ws];
            if (wx) v += ax * wx[(long)k * ws];
            v += -beta * (res[k * T] - src[(long)k * ws]) * dt;
            wn[k] = v;
            Wn[(long)k * m.sC + c] = v;
        }
        if (Qn) { Prim<R> q; primitive(ph, wn[0], wn + 1, wn[4], q); store_prim(Qn, m.sN, c, q); }
        return res[5 * T];
    }

#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        std::vector<R> qg((size_t)20 * TS, R(0)), vol(T, R(1)), acc((size_t)6 * T, R(0));
        for (int l = 0; l < nc; l++) { tile_stage_cell<R, TS>(qg.data(), l, Q, G, m.sN, c0 + l); vol[l] = m.vol[c0 + l]; }
        for (int h = m.halo_start[t]; h < m.halo_start[t + 1]; h++) tile_stage_cell<R, TS>(qg.data(), T + h - m.halo_start[t], Q, G, m.sN, m.halo_cell[h]);
        for (int p = m.pass_start[t]; p < m.pass_start[t + 1]; p++) {
            const R* chunk = m.chunks + (long)p * Chunk<R, W>::kScalars;
            for (int i = 0; i < W; i++) {
                TileEntry e; tile_decode(Chunk<R, W>::words(chunk)[i], e);
                if (!e.valid) continue;
                if ((e.lo >= T || e.ln >= T) && p - m.pass_start[t] < m.halo_pass[t]) throw std::runtime_error("halo_pass inconsistent");
                Geom<R> gm; tile_load_geom<R, W>(chunk, i, gm);
                Flux5<R> F; R wave, sO, sNb;
                face(gm, e, qg.data(), vol.data(), F, wave, sO, sNb);
                scatter(acc.data(), e.lo, e.ln, F, wave, sO, sNb);
            }
        }
        const R* wx = W1 ? W1 : W2; const R ax = W1 ? a1 : a2;
        R mx = R(-1e30);
        for (int l = 0; l < nc; l++) { const int c = c0 + l; R d = finish(c, &acc[l], W0 + c, wx ? wx + c : nullptr, ax, S + c, m.sC); mx = d > mx ? d : mx; }
        if (dtc_partial) dtc_partial[t] = mx;
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x;
        R* qg = reinterpret_cast<R*>(smem);
        R* vol = qg + 20 * TS;
        R* acc = vol + T;
        R* chunk = reinterpret_cast<R*>(smem + Smem::kChunkOff);
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + Smem::kBarOff);
        const int p0 = m.pass_start[t], np = m.pass_start[t + 1] - p0;
        const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0, hp = m.halo_pass[t];
        const R* wx = W1 ? W1 : W2; const R ax = W1 ? a1 : a2;
        constexpr unsigned kRow = T * (unsigned)sizeof(R);
        if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_fence_init(); }
        __syncthreads();
        if (tid == 0) {
            // the tile's own rows (may run past the last internal cell into ghost rows / slack: allocated, unused) + first chunk
            mbar_expect_tx(&bars[0], 21u * kRow);
            for (int k = 0; k < 5; k++) bulk_g2s(qg + k * TS, Q + (long)k * m.sN + c0, kRow, &bars[0]);
            for (int k = 0; k < 15; k++) bulk_g2s(qg + (5 + k) * TS, G + (long)k * m.sN + c0, kRow, &bars[0]);
            bulk_g2s(vol, m.vol + c0, kRow, &bars[0]);
            mbar_expect_tx(&bars[1], (unsigned)Chunk<R, W>::kBytes);
            bulk_g2s(chunk, m.chunks + (long)p0 * Chunk<R, W>::kScalars, (unsigned)Chunk<R, W>::kBytes, &bars[1]);
        }
        // halo rows: asynchronous gather, consumed from pass `hp` on (the passes before it only touch the tile's own cells)
        for (int h = tid; h < nh; h += W) {
            const int cell = m.halo_cell[h0 + h];
            for (int k = 0; k < 5; k++) cp_async_elem<sizeof(R)>(qg + k * TS + T + h, Q + (long)k * m.sN + cell);
            for (int k = 0; k < 15; k++) cp_async_elem<sizeof(R)>(qg + (5 + k) * TS + T + h, G + (long)k * m.sN + cell);
        }
        cp_async_commit();
        for (int i = tid; i < 6 * T; i += W) acc[i] = R(0);
        mbar_wait(&bars[0], 0);
        for (int p = 0; p < np; p++) {
            mbar_wait(&bars[1], (unsigned)(p & 1));
            TileEntry e; tile_decode(Chunk<R, W>::words(chunk)[tid], e);
            Geom<R> gm; tile_load_geom<R, W>(chunk, tid, gm);
            const int cfirst = (int)((Chunk<R, W>::words(chunk)[0] >> 20) & 0x1Fu), clast = (int)((Chunk<R, W>::words(chunk)[W - 1] >> 20) & 0x1Fu);
            if (p == hp) cp_async_wait_all();
            __syncthreads();                       // metrics are in registers: the chunk buffer is free; halo rows visible from pass hp on
            if (tid == 0) {
                fence_proxy_async();
                if (p + 1 < np) {
                    mbar_expect_tx(&bars[1], (unsigned)Chunk<R, W>::kBytes);
                    bulk_g2s(chunk, m.chunks + (long)(p0 + p + 1) * Chunk<R, W>::kScalars, (unsigned)Chunk<R, W>::kBytes, &bars[1]);
                } else {
                    // last pass: the chunk buffer now receives the RK update's inputs (W0, S, W1|W2 rows of the tile)
                    mbar_expect_tx(&bars[2], (wx ? 15u : 10u) * kRow);
                    for (int k = 0; k < 5; k++) {
                        bulk_g2s(chunk + k * T, W0 + (long)k * m.sC + c0, kRow, &bars[2]);
                        bulk_g2s(chunk + (5 + k) * T, S + (long)k * m.sC + c0, kRow, &bars[2]);
                        if (wx) bulk_g2s(chunk + (10 + k) * T, wx + (long)k * m.sC + c0, kRow, &bars[2]);
                    }
                }
            }
            Flux5<R> F; R wave = R(0), sO = R(0), sNb = R(0);
            if (e.valid) face(gm, e, qg, vol, F, wave, sO, sNb);
            const int col = e.valid ? e.col : -1;
            for (int c = cfirst; c <= clast; c++) {
                if (col == c) scatter(acc, e.lo, e.ln, F, wave, sO, sNb);
                __syncthreads();
            }
        }
        mbar_wait(&bars[2], 0);
        R mx = R(-1e30);
        for (int l = tid; l < nc; l += W) { R d = finish(c0 + l, acc + l, chunk + l, wx ? chunk + 10 * T + l : nullptr, ax, chunk + 5 * T + l, T); mx = d > mx ? d : mx; }
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
    static constexpr int kMinBlocks = 1;
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G; const R* abar; R coef;
    R *Qb, *Gb;
#if defined(__CUDACC__)
    typedef TileSmem<R, T, TS, W, 6 * TS + 20 * T> Smem;
    static size_t smem_bytes() { return Smem::kBytes; }
#endif

    // ab [5][TS]: abar of the slot's cell (0 for ghost cells: boundary faces scatter to their owner only), vol [TS]
    FVM_HD void stage_ab(R* ab, R* vol, int slot, int cell) const {
        if (cell < m.nInternalCells) { for (int k = 0; k < 5; k++) ab[k * TS + slot] = abar[(long)k * m.sC + cell]; vol[slot] = m.vol[cell]; }
        else { for (int k = 0; k < 5; k++) ab[k * TS + slot] = R(0); vol[slot] = R(1); }
    }
    FVM_HD void face(const Geom<R>& gm, const TileEntry& e, const R* qg, const R* ab, const R* vol, Prim<R>& qLb, Grad<R>& gLb, Prim<R>& qRb, Grad<R>& gRb) const {
        Prim<R> qL, qR; Grad<R> gL, gR;
        tile_load_cell<R, TS>(qg, e.lo, qL, gL); tile_load_cell<R, TS>(qg, e.ln, qR, gR);
        const R sO = gm.area * coef * rcp(vol[e.lo]), sN_ = gm.area * coef * rcp(vol[e.ln]);
        R d[5];
        for (int k = 0; k < 5; k++) d[k] = ab[k * TS + e.lo] * sO - ab[k * TS + e.ln] * sN_;
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
        std::vector<R> qg((size_t)20 * TS, R(0)), ab((size_t)5 * TS, R(0)), vol(TS, R(1)), acc((size_t)20 * T, R(0));
        for (int l = 0; l < nc; l++) { tile_stage_cell<R, TS>(qg.data(), l, Q, G, m.sN, c0 + l); stage_ab(ab.data(), vol.data(), l, c0 + l); }
        for (int h = m.halo_start[t]; h < m.halo_start[t + 1]; h++) {
            const int slot = T + h - m.halo_start[t];
            tile_stage_cell<R, TS>(qg.data(), slot, Q, G, m.sN, m.halo_cell[h]); stage_ab(ab.data(), vol.data(), slot, m.halo_cell[h]);
        }
        for (int p = m.pass_start[t]; p < m.pass_start[t + 1]; p++) {
            const R* chunk = m.chunks + (long)p * Chunk<R, W>::kScalars;
            for (int i = 0; i < W; i++) {
                TileEntry e; tile_decode(Chunk<R, W>::words(chunk)[i], e);
                if (!e.valid) continue;
                Geom<R> gm; tile_load_geom<R, W>(chunk, i, gm);
                Prim<R> qLb, qRb; Grad<R> gLb, gRb;
                face(gm, e, qg.data(), ab.data(), vol.data(), qLb, gLb, qRb, gRb);
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
        R* ab = qg + 20 * TS;
        R* vol = ab + 5 * TS;
        R* acc = vol + TS;
        R* chunk = reinterpret_cast<R*>(smem + Smem::kChunkOff);
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + Smem::kBarOff);
        const int p0 = m.pass_start[t], np = m.pass_start[t + 1] - p0;
        const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0, hp = m.halo_pass[t];
        constexpr unsigned kRow = T * (unsigned)sizeof(R);
        if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bars[0], 26u * kRow);
            for (int k = 0; k < 5; k++) bulk_g2s(qg + k * TS, Q + (long)k * m.sN + c0, kRow, &bars[0]);
            for (int k = 0; k < 15; k++) bulk_g2s(qg + (5 + k) * TS, G + (long)k * m.sN + c0, kRow, &bars[0]);
            for (int k = 0; k < 5; k++) bulk_g2s(ab + k * TS, abar + (long)k * m.sC + c0, kRow, &bars[0]);
            bulk_g2s(vol, m.vol + c0, kRow, &bars[0]);
            mbar_expect_tx(&bars[1], (unsigned)Chunk<R, W>::kBytes);
            bulk_g2s(chunk, m.chunks + (long)p0 * Chunk<R, W>::kScalars, (unsigned)Chunk<R, W>::kBytes, &bars[1]);
        }
        for (int h = tid; h < nh; h += W) {
            const int cell = m.halo_cell[h0 + h], slot = T + h;
            for (int k = 0; k < 5; k++) cp_async_elem<sizeof(R)>(qg + k * TS + slot, Q + (long)k * m.sN + cell);
            for (int k = 0; k < 15; k++) cp_async_elem<sizeof(R)>(qg + (5 + k) * TS + slot, G + (long)k * m.sN + cell);
            if (cell < m.nInternalCells) {
                for (int k = 0; k < 5; k++) cp_async_elem<sizeof(R)>(ab + k * TS + slot, abar + (long)k * m.sC + cell);
                cp_async_elem<sizeof(R)>(vol + slot, m.vol + cell);
            } else {
                for (int k = 0; k < 5; k++) ab[k * TS + slot] = R(0);
                vol[slot] = R(1);
            }
        }
        cp_async_commit();
        for (int i = tid; i < 20 * T; i += W) acc[i] = R(0);
        mbar_wait(&bars[0], 0);
        for (int p = 0; p < np; p++) {
            mbar_wait(&bars[1], (unsigned)(p & 1));
            TileEntry e; tile_decode(Chunk<R, W>::words(chunk)[tid], e);
            Geom<R> gm; tile_load_geom<R, W>(chunk, tid, gm);
            const int cfirst = (int)((Chunk<R, W>::words(chunk)[0] >> 20) & 0x1Fu), clast = (int)((Chunk<R, W>::words(chunk)[W - 1] >> 20) & 0x1Fu);
            if (p == hp) cp_async_wait_all();
            __syncthreads();
            if (tid == 0 && p + 1 < np) {
                fence_proxy_async();
                mbar_expect_tx(&bars[1], (unsigned)Chunk<R, W>::kBytes);
                bulk_g2s(chunk, m.chunks + (long)(p0 + p + 1) * Chunk<R, W>::kScalars, (unsigned)Chunk<R, W>::kBytes, &bars[1]);
            }
            Prim<R> qLb, qRb; Grad<R> gLb, gRb;
            int ghost = -1;
            if (e.valid) { face(gm, e, qg, ab, vol, qLb, gLb, qRb, gRb); ghost = ghost_of(t, e); }
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
