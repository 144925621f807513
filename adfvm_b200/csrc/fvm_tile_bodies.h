// Tiled flux kernels: one CTA per tile of T consecutive cells (fvm_tiles.h), one WARP per sub-tile of 32 cells,
// one LANE per cell: the lane evaluates the entries homed at its cell (one per round, always as the owner side,
// fvm_tiles.h) and keeps the cell's accumulators in registers.
//
//   flux_tile       a8-a12: face flux (reconstruction + Riemann + viscous, adFVM/density.py:253-331) of every face
//                   touching the sub-tile, +F*A/V to the home cell, -F*A/V_other sent to the other cell's lane by warp
//                   shuffle (adFVM/op.py:12-29), then RK stage update (adFVM/timestep.py:35-45) and the primitive
//                   conversion of the new state (adFVM/density.py:162-171) of the lane's cell.
//   flux_grad_tile  reverse of the flux + scatter: VJP of every face of the sub-tile in compact form (FaceAdj), the home
//                   cell's 20 input adjoints summed in registers, the other cell's sent by shuffle; ghost rows of
//                   boundary faces written directly (exclusive writer).
//
// Data movement of one CTA (R = scalar):
//   * qg [20][TS]: U(3),T,p and the 15 gradient components of the tile's own cells (slots [0,T): 20 bulk row copies of
//     the SoA arrays, cp.async.bulk / SASS UBLKCP, completion on an mbarrier) and of its halo (slots [T,T+nHalo): cells
//     of other tiles / ghost cells, gathered once with LDGSTS, completion on a second mbarrier);
//   * face metrics live in HBM as per-round CHUNKS [16][32] R + [32] u32 (area, n, 1/delta, dUnit, linW, quadW and the
//     packed entry word) in the order the warp consumes them; one bulk copy per warp and round streams a chunk into the
//     warp's own buffer, issued one round ahead of its use, completion on the warp's own mbarrier;
//   * forward:  vol [T]          reverse:  ab [5][TS] = abar (own slots scaled by 1/V once staged), volh [TS-T] = V of halo slots.
// After the staging the warps of a CTA never synchronise with each other. The CPU simulator (tests/hostsim) runs the
// same face / share functions round by round in the same order.
#pragma once
#include "fvm_bodies.h"
#if !defined(__CUDACC__)
#include <limits>
#include <stdexcept>
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

struct TileEntry { int ln, kind, src; bool valid, sn, ghost, has_src; };   // sn: the other cell's share is sent to its lane
FVM_HD void tile_decode(unsigned w, TileEntry& e) {
    e.src = (int)(w & 0x1Fu); e.has_src = ((w >> 5) & 1u) != 0; e.ln = (int)((w >> 10) & 0x3FFu);
    e.kind = (int)((w >> 25) & 3u); e.valid = ((w >> 27) & 1u) != 0;
    e.sn = ((w >> 29) & 1u) != 0; e.ghost = ((w >> 30) & 1u) != 0;
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
            if (slot_word[i] >> 31) {        // home cell = the face's neighbour: reverse the face (fvm_tiles.h)
                for (int k = 0; k < 3; k++) { v[1 + k] = -v[1 + k]; v[5 + k] = -v[5 + k]; const R q = v[10 + k]; v[10 + k] = v[13 + k]; v[13 + k] = q; }
                const R l = v[8]; v[8] = v[9]; v[9] = l;
            }
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
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename R> __device__ __forceinline__ R warp_max(R v) {
    for (int o = 16; o > 0; o >>= 1) { R x = __shfl_down_sync(0xffffffffu, v, o); v = x > v ? x : v; }
    return v;   // valid in lane 0
}

// Ampere-style asynchronous element copies (SASS LDGSTS) for the halo gather: no registers, no stall at issue
template <int BYTES> __device__ __forceinline__ void cp_async_elem(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
// the mbarrier (initialised with one pending arrival per thread) receives this thread's arrival once all its earlier
// cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// L2 prefetch of the index data a later tile needs before its first round. i = job of the calling lane:
//   0..6      halo cell list of tile tn (up to 7 lines of 128 B)          } tn runs one wave of CTAs after this tile; its
//   7..6+NW   first metric chunk of each warp of tile tn                   } index lines were prefetched one wave earlier,
//   7+NW..    the index lines (round_start, halo_start, halo_round) of tile tnn, two waves ahead
template <typename R, int NW, class Ch> __device__ __forceinline__ void tile_prefetch_meta(const MeshDev<R>& m, int tn, int tnn, int i) {
    if (i < 7) {
        const int hs = m.halo_start[tn], he = m.halo_start[tn + 1];
        const int* p = m.halo_cell + hs + 32 * i;
        if (p < m.halo_cell + he) prefetch_l2(p);
    } else if (i < 7 + NW) {
        const int r = m.round_start[tn * NW + (i - 7)];
        if (r < m.round_start[tn * NW + (i - 7) + 1]) bulk_prefetch_l2(m.chunks + (long)r * Ch::kScalars, (unsigned)Ch::kBytes);
    } else if (tnn < m.nTiles) {
        if (i == 7 + NW) prefetch_l2(m.round_start + tnn * NW);
        else if (i == 8 + NW) prefetch_l2(m.halo_start + tnn);
        else if (i == 9 + NW) prefetch_l2(m.halo_round + tnn * NW);
    }
}

// shared-memory carve-up common to both kernels: [qg 20*TS][extra ...][one chunk per warp][2 + T/32 mbarriers]
template <typename R, int T, int TS, int EXTRA> struct TileSmem {
    static constexpr int NW = T / 32;
    static constexpr size_t kChunkOff = ((size_t)(20 * TS + EXTRA) * sizeof(R) + 127) / 128 * 128;
    static constexpr size_t kBarOff = kChunkOff + (size_t)NW * Chunk<R, 32>::kBytes;
    static constexpr size_t kBytes = kBarOff + 8 * (2 + NW);
};
#endif

// ------------------------------------------------------------------------------------------ forward
template <typename R, int T, int TS> struct FluxTileBody {
    static constexpr const char* kName = "flux_tile";
    static constexpr int kThreads = T, NW = T / 32;
    static constexpr int kMinBlocks = sizeof(R) == 8 ? 3 : 5;     // register cap: 168 (fp64), 96 (fp32)
    static constexpr int kPrefetchDistance = 148 * kMinBlocks;    // tiles per wave of resident CTAs
    typedef Chunk<R, 32> Ch;
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G;            // this stage's primitives / gradients (ghosts filled)
    const R *W0, *W1, *W2;     // previous stage states (W1/W2 NULL when their alpha is 0; never both set in SSPRK3)
    R a0, a1, a2, beta, dt;
    const R* S;                // source terms [5][sC]
    R* Wn;                     // new state
    R* Qn;                     // primitives of the new state (may be NULL)
    R* dtc_partial;            // [nTiles*NW] per-sub-tile max of dtc (may be NULL)
#if defined(__CUDACC__)
    typedef TileSmem<R, T, TS, T> Smem;
    static size_t smem_bytes() { return Smem::kBytes; }
#endif

    // one entry of the lane whose cell sits in slot lo: own[6] = the home cell's share of (residual(5), dtc),
    // snd[6] = the other cell's share (zero unless e.sn). ivh = 1/V_home, vol: the tile's own volumes [T]
    FVM_HD void face(const Geom<R>& gm, const TileEntry& e, int lo, const R* qg, R ivh, const R* vol, R* own, R* snd) const {
        Prim<R> qL, qR; Grad<R> gL, gR;
        tile_load_cell<R, TS>(qg, lo, qL, gL); tile_load_cell<R, TS>(qg, e.ln, qR, gR);
        Flux5<R> F; R wave;
        face_flux(ph, e.kind, gm, qL, gL, qR, gR, F, wave);
        const R sO = gm.area * ivh;
        const R sNb = e.sn ? gm.area * rcp(vol[e.ln]) : R(0);
        own[0] = F.rho * sO; own[1] = F.rhoU[0] * sO; own[2] = F.rhoU[1] * sO; own[3] = F.rhoU[2] * sO; own[4] = F.rhoE * sO; own[5] = wave * sO;
        const R s = -sNb;
        snd[0] = F.rho * s; snd[1] = F.rhoU[0] * s; snd[2] = F.rhoU[1] * s; snd[3] = F.rhoU[2] * s; snd[4] = F.rhoE * s; snd[5] = wave * sNb;
    }
    // RK stage update + primitives of the new state for cell c; res[0..4] = residual, res[5] = dtc;
    // w0/wx/src: this cell's previous-stage states and source, component stride ws (wx = W1 or W2, coefficient ax)
    FVM_HD R finish(int c, const R* res, const R* w0, const R* wx, R ax, const R* src, int ws) const {
        R wn[5];
        for (int k = 0; k < 5; k++) {
            R v = a0 * w0[(long)k * ws];
            if (wx) v += ax * wx[(long)k * ws];
            v += -beta * (res[k] - src[(long)k * ws]) * dt;
            wn[k] = v;
            Wn[(long)k * m.sC + c] = v;
        }
        if (Qn) { Prim<R> q; primitive(ph, wn[0], wn + 1, wn[4], q); store_prim(Qn, m.sN, c, q); }
        return res[5];
    }

#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        std::vector<R> qg((size_t)20 * TS, R(0)), vol(T, R(1));
        for (int l = 0; l < nc; l++) { tile_stage_cell<R, TS>(qg.data(), l, Q, G, m.sN, c0 + l); vol[l] = m.vol[c0 + l]; }
        for (int h = m.halo_start[t]; h < m.halo_start[t + 1]; h++) tile_stage_cell<R, TS>(qg.data(), T + h - m.halo_start[t], Q, G, m.sN, m.halo_cell[h]);
        const R* wx = W1 ? W1 : W2; const R ax = W1 ? a1 : a2;
        for (int w = 0; w < NW; w++) {
            const int r0 = m.round_start[t * NW + w], r1 = m.round_start[t * NW + w + 1];
            R acc[32][6], snd[32][6], own[6];
            for (int l = 0; l < 32; l++) for (int k = 0; k < 6; k++) acc[l][k] = R(0);
            for (int r = r0; r < r1; r++) {
                const R* chunk = m.chunks + (long)r * Ch::kScalars;
                TileEntry e[32];
                for (int i = 0; i < 32; i++) {
                    tile_decode(Ch::words(chunk)[i], e[i]);
                    for (int k = 0; k < 6; k++) snd[i][k] = R(0);
                    if (!e[i].valid) continue;
                    if (e[i].ln >= T && r - r0 < m.halo_round[t * NW + w]) throw std::runtime_error("halo_round inconsistent");
                    if (w * 32 + i >= nc) throw std::runtime_error("entry homed at a lane without a cell");
                    if (e[i].sn && (e[i].ln >> 5) != w) throw std::runtime_error("share sent outside the sub-tile");
                    Geom<R> gm; tile_load_geom<R, 32>(chunk, i, gm);
                    face(gm, e[i], w * 32 + i, qg.data(), R(1) / vol[w * 32 + i], vol.data(), own, snd[i]);
                    for (int k = 0; k < 6; k++) acc[i][k] += own[k];
                }
                for (int i = 0; i < 32; i++) {
                    if (!e[i].has_src) continue;
                    const TileEntry& s = e[e[i].src];
                    if (!s.valid || !s.sn || s.ln != w * 32 + i) throw std::runtime_error("receive lane inconsistent");
                    for (int k = 0; k < 6; k++) acc[i][k] += snd[e[i].src][k];
                }
                for (int i = 0; i < 32; i++) if (e[i].valid && e[i].sn) {      // every sent share has exactly one receiver
                    const TileEntry& d = e[e[i].ln - w * 32];
                    if (!d.has_src || d.src != i) throw std::runtime_error("sent share without receiver");
                }
            }
            R mx = R(-1e30);
            for (int l = w * 32; l < nc && l < w * 32 + 32; l++) {
                const int c = c0 + l;
                R d = finish(c, acc[l - w * 32], W0 + c, wx ? wx + c : nullptr, ax, S + c, m.sC); mx = d > mx ? d : mx;
            }
            if (dtc_partial) dtc_partial[t * NW + w] = mx;
        }
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
        R* qg = reinterpret_cast<R*>(smem);
        R* vol = qg + 20 * TS;
        R* chunk = reinterpret_cast<R*>(smem + Smem::kChunkOff) + w * Ch::kScalars;
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + Smem::kBarOff);
        unsigned long long *bar_rows = &bars[0], *bar_halo = &bars[1], *bar_chunk = &bars[2 + w];
        const int r0 = m.round_start[t * NW + w], nr = m.round_start[t * NW + w + 1] - r0;
        const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0, hr = m.halo_round[t * NW + w];
        const R* wx = W1 ? W1 : W2; const R ax = W1 ? a1 : a2;
        constexpr unsigned kRow = T * (unsigned)sizeof(R), kSub = 32u * (unsigned)sizeof(R);
        if (tid == 0) {
            mbar_init(bar_rows, 1); mbar_init(bar_halo, T);
            for (int i = 0; i < NW; i++) mbar_init(&bars[2 + i], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            // the tile's own rows (may run past the last internal cell into ghost rows / slack: allocated, unused)
            mbar_expect_tx(bar_rows, 21u * kRow);
            for (int k = 0; k < 5; k++) bulk_g2s(qg + k * TS, Q + (long)k * m.sN + c0, kRow, bar_rows);
            for (int k = 0; k < 15; k++) bulk_g2s(qg + (5 + k) * TS, G + (long)k * m.sN + c0, kRow, bar_rows);
            bulk_g2s(vol, m.vol + c0, kRow, bar_rows);
        }
        if (lane == 0 && nr > 0) {
            mbar_expect_tx(bar_chunk, (unsigned)Ch::kBytes);
            bulk_g2s(chunk, m.chunks + (long)r0 * Ch::kScalars, (unsigned)Ch::kBytes, bar_chunk);
        }
        // halo rows: asynchronous gather, consumed from round `hr` on (the rounds before it only touch the tile's own cells)
        {
            constexpr int kIter = (TS - T + T - 1) / T;          // halo slots per thread
            int hc[kIter];
            #pragma unroll
            for (int i = 0; i < kIter; i++) { const int h = tid + i * T; hc[i] = h < nh ? m.halo_cell[h0 + h] : -1; }   // all index loads in flight at once
            #pragma unroll
            for (int i = 0; i < kIter; i++) {
                const int h = tid + i * T, cell = hc[i];
                if (cell < 0) continue;
                for (int k = 0; k < 5; k++) cp_async_elem<sizeof(R)>(qg + k * TS + T + h, Q + (long)k * m.sN + cell);
                for (int k = 0; k < 15; k++) cp_async_elem<sizeof(R)>(qg + (5 + k) * TS + T + h, G + (long)k * m.sN + cell);
            }
        }
        cp_async_arrive(bar_halo);
        // pull the rows of the tile that will run in this CTA slot one wave later into L2: its first loads then see an
        // L2 hit instead of a loaded-DRAM round trip (the CTA's start-up latency is exposed at 8-12 warps per SM)
        {
            const int tn = t + kPrefetchDistance;
            if (tn < m.nTiles && w == NW - 1) {
                const long cn = (long)tn * T;
                if (lane < 5) bulk_prefetch_l2(Q + (long)lane * m.sN + cn, kRow);
                else if (lane < 20) bulk_prefetch_l2(G + (long)(lane - 5) * m.sN + cn, kRow);
                else if (lane == 20) bulk_prefetch_l2(m.vol + cn, kRow);
            }
            if (tn < m.nTiles && w == (NW > 1 ? NW - 2 : 0)) tile_prefetch_meta<R, NW, Ch>(m, tn, tn + kPrefetchDistance, lane);
        }
        if (nr == 0) return;                       // sub-tile beyond the last internal cell
        mbar_wait(bar_rows, 0);
        const R ivh = tid < nc ? rcp(vol[tid]) : R(0);
        R acc[6] = {R(0), R(0), R(0), R(0), R(0), R(0)};
        for (int r = 0; r < nr; r++) {
            mbar_wait(bar_chunk, (unsigned)(r & 1));
            TileEntry e; tile_decode(Ch::words(chunk)[lane], e);
            Geom<R> gm; tile_load_geom<R, 32>(chunk, lane, gm);
            if (r == hr) mbar_wait(bar_halo, 0);
            __syncwarp();                          // metrics are in registers: the warp's chunk buffer is free
            if (lane == 0) {
                fence_proxy_async();
                if (r + 1 < nr) {
                    mbar_expect_tx(bar_chunk, (unsigned)Ch::kBytes);
                    bulk_g2s(chunk, m.chunks + (long)(r0 + r + 1) * Ch::kScalars, (unsigned)Ch::kBytes, bar_chunk);
                } else {
                    // last round: the chunk buffer now receives the RK update's inputs (W0, S, W1|W2 of the sub-tile's cells)
                    const long cw = c0 + w * 32;
                    mbar_expect_tx(bar_chunk, (wx ? 15u : 10u) * kSub);
                    for (int k = 0; k < 5; k++) {
                        bulk_g2s(chunk + k * 32, W0 + (long)k * m.sC + cw, kSub, bar_chunk);
                        bulk_g2s(chunk + (5 + k) * 32, S + (long)k * m.sC + cw, kSub, bar_chunk);
                        if (wx) bulk_g2s(chunk + (10 + k) * 32, wx + (long)k * m.sC + cw, kSub, bar_chunk);
                    }
                }
            }
            R own[6] = {R(0), R(0), R(0), R(0), R(0), R(0)}, snd[6] = {R(0), R(0), R(0), R(0), R(0), R(0)};
            if (e.valid) face(gm, e, tid, qg, ivh, vol, own, snd);
            for (int k = 0; k < 6; k++) acc[k] += own[k];
            if (__any_sync(0xffffffffu, e.has_src)) {
                for (int k = 0; k < 6; k++) { const R v = __shfl_sync(0xffffffffu, snd[k], e.src); if (e.has_src) acc[k] += v; }
            }
        }
        mbar_wait(bar_chunk, (unsigned)(nr & 1));
        R mx = R(-1e30);
        if (tid < nc) mx = finish(c0 + tid, acc, chunk + lane, wx ? chunk + 10 * 32 + lane : nullptr, ax, chunk + 5 * 32 + lane, 32);
        if (dtc_partial) {
            mx = warp_max(mx);
            if (lane == 0) dtc_partial[t * NW + w] = mx;
        }
    }
#endif
};

// ------------------------------------------------------------------------------------------ reverse
//   abar = adjoint of the stage OUTPUT state [5][sC]; coef = -beta_ii*dt (d W_new / d residual)
//   outputs Qb [5][sN], Gb [15][sN]: rows of internal cells (Gb divided by the cell volume, see GradAdjUpdateBody)
//   and of the ghost cells of ALL boundary faces
// SPEC: 0 = Riemann solver and viscosity law read from Phys at run time; 1 = Roe + Sutherland, 2 = Roe + constant viscosity as
// compile-time constants (the device code works on a copy of Phys with those fields overwritten: the branches on them fold, the
// coupled-face VJP loses its dead Lax-Friedrichs / other-law code; measured -6 % on the kernel at 256^3, profiles/README.md)
enum { SPEC_GENERIC = 0, SPEC_ROE_SUTHERLAND = 1, SPEC_ROE_CONSTANT = 2 };
template <typename R, int T, int TS, int SPEC = 0> struct FluxGradTileBody {
    static constexpr const char* kName = "flux_grad_tile";
    static constexpr int kThreads = T, NW = T / 32;
    // fp64: 254 registers, 2 CTAs per SM. A third CTA (168 registers; the compact variant fits three in shared memory) was
    // measured and is slower: 400 B of spills per thread, 1.20 ms instead of 0.76 ms at 128^3 (profiles/README.md, round 2)
#ifndef ADFVM_FG_MINBLOCKS
#define ADFVM_FG_MINBLOCKS 2
#endif
    static constexpr int kMinBlocks = sizeof(R) == 8 ? ADFVM_FG_MINBLOCKS : 4;
    static constexpr int kPrefetchDistance = 148 * kMinBlocks;
    typedef Chunk<R, 32> Ch;
    Phys<R> ph; MeshDev<R> m;
    const R *Q, *G; const R* abar; R coef;
    R *Qb, *Gb;
#if defined(__CUDACC__)
    typedef TileSmem<R, T, TS, 5 * TS + (TS - T)> Smem;
    static size_t smem_bytes() { return Smem::kBytes; }
#endif

    // ab [5][TS]: abar of the slot's cell - slots [0,T) (the tile's own cells) already divided by the cell volume, halo slots raw
    // with their volume in volh [TS-T]; ghost slots are never read (boundary faces scatter to their owner only)
    FVM_HD void stage_ab(R* ab, R* volh, int slot, int cell) const {
        if (cell >= m.nInternalCells) return;
        if (slot < T) { const R iv = R(1) / m.vol[cell]; for (int k = 0; k < 5; k++) ab[k * TS + slot] = abar[(long)k * m.sC + cell] * iv; }
        else { for (int k = 0; k < 5; k++) ab[k * TS + slot] = abar[(long)k * m.sC + cell]; volh[slot - T] = m.vol[cell]; }
    }
    // one entry of the lane whose cell sits in slot lo: own[20] += the home cell's input adjoints; snd[20] = the other
    // cell's (set when e.sn; zero otherwise); ghost >= 0: global row of the other cell when it is a ghost cell, whose
    // adjoints are stored straight to Qb/Gb
    FVM_HD void face(const Phys<R>& ph, const Geom<R>& gm, const TileEntry& e, int lo, int ghost, const R* qg, const R* ab, const R* volh, R* own, R* snd) const {
        Prim<R> qL, qR; Grad<R> gL, gR;
        tile_load_cell<R, TS>(qg, lo, qL, gL); tile_load_cell<R, TS>(qg, e.ln, qR, gR);
        const R sO = gm.area * coef;
        R d[5];
        for (int k = 0; k < 5; k++) d[k] = ab[k * TS + lo] * sO;
        if (!e.ghost) {
            const R sN_ = e.ln >= T ? sO * rcp(volh[e.ln - T]) : sO;
            for (int k = 0; k < 5; k++) d[k] -= ab[k * TS + e.ln] * sN_;
        }
        Flux5<R> Fb; Fb.rho = d[0]; Fb.rhoU[0] = d[1]; Fb.rhoU[1] = d[2]; Fb.rhoU[2] = d[3]; Fb.rhoE = d[4];
        if (e.kind == FACE_COUPLED) {
            FaceAdj<R> c; face_flux_vjp_coupled(ph, gm, qL, gL, qR, gR, Fb, c);
            for (int k = 0; k < 20; k++) own[k] += face_adj_owner(c, gm, k);
            if (ghost >= 0) { for (int k = 0; k < 20; k++) (k < 5 ? Qb : Gb)[(long)(k < 5 ? k : k - 5) * m.sN + ghost] = face_adj_neighbour(c, gm, k); }
            else if (e.sn) for (int k = 0; k < 20; k++) snd[k] = face_adj_neighbour(c, gm, k);
        } else {
            Prim<R> qLb, qRb; Grad<R> gLb, gRb;
            zero(qLb); zero(gLb); zero(qRb); zero(gRb);
            face_flux_vjp(ph, e.kind, gm, qL, gL, qR, gR, Fb, qLb, gLb, qRb, gRb);
            own[0] += qLb.U[0]; own[1] += qLb.U[1]; own[2] += qLb.U[2]; own[3] += qLb.T; own[4] += qLb.p;
            for (int k = 0; k < 9; k++) own[5 + k] += gLb.U[k];
            for (int k = 0; k < 3; k++) { own[14 + k] += gLb.T[k]; own[17 + k] += gLb.p[k]; }
            if (ghost >= 0) { store_prim(Qb, m.sN, ghost, qRb); store_grad(Gb, m.sN, ghost, gRb); }
        }
    }
    FVM_HD int ghost_of(int t, const TileEntry& e) const { return e.ghost ? m.halo_cell[m.halo_start[t] + e.ln - T] : -1; }
    // internal rows of Gb are stored divided by the cell volume (what GradAdjUpdateBody consumes); iv = 1/V_c
    FVM_HD void finish(int c, const R* a, R iv) const {
        for (int k = 0; k < 5; k++) Qb[(long)k * m.sN + c] = a[k];
        for (int k = 0; k < 15; k++) Gb[(long)k * m.sN + c] = a[5 + k] * iv;
    }

#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        const R nan = std::numeric_limits<R>::quiet_NaN();       // ghost slots of ab / volh must never be read
        std::vector<R> qg((size_t)20 * TS, R(0)), ab((size_t)5 * TS, nan), volh(TS - T, nan);
        for (int l = 0; l < nc; l++) { tile_stage_cell<R, TS>(qg.data(), l, Q, G, m.sN, c0 + l); stage_ab(ab.data(), volh.data(), l, c0 + l); }
        for (int h = m.halo_start[t]; h < m.halo_start[t + 1]; h++) {
            const int slot = T + h - m.halo_start[t];
            tile_stage_cell<R, TS>(qg.data(), slot, Q, G, m.sN, m.halo_cell[h]); stage_ab(ab.data(), volh.data(), slot, m.halo_cell[h]);
        }
        for (int w = 0; w < NW; w++) {
            R acc[32][20], snd[32][20];
            for (int l = 0; l < 32; l++) for (int k = 0; k < 20; k++) acc[l][k] = R(0);
            for (int r = m.round_start[t * NW + w]; r < m.round_start[t * NW + w + 1]; r++) {
                const R* chunk = m.chunks + (long)r * Ch::kScalars;
                TileEntry e[32];
                for (int i = 0; i < 32; i++) {
                    tile_decode(Ch::words(chunk)[i], e[i]);
                    for (int k = 0; k < 20; k++) snd[i][k] = R(0);
                    if (!e[i].valid) continue;
                    if (e[i].ghost != (e[i].ln >= T && m.halo_cell[m.halo_start[t] + e[i].ln - T] >= m.nInternalCells)) throw std::runtime_error("ghost flag inconsistent");
                    if (e[i].kind != FACE_COUPLED && !e[i].ghost) throw std::runtime_error("boundary-kind entry without a ghost cell");
                    Geom<R> gm; tile_load_geom<R, 32>(chunk, i, gm);
                    face(ph, gm, e[i], w * 32 + i, ghost_of(t, e[i]), qg.data(), ab.data(), volh.data(), acc[i], snd[i]);
                }
                for (int i = 0; i < 32; i++) if (e[i].has_src) for (int k = 0; k < 20; k++) acc[i][k] += snd[e[i].src][k];
            }
            for (int l = w * 32; l < nc && l < w * 32 + 32; l++) finish(c0 + l, acc[l - w * 32], R(1) / m.vol[c0 + l]);
        }
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
        R* qg = reinterpret_cast<R*>(smem);
        R* ab = qg + 20 * TS;
        R* volh = ab + 5 * TS;
        R* chunk = reinterpret_cast<R*>(smem + Smem::kChunkOff) + w * Ch::kScalars;
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + Smem::kBarOff);
        unsigned long long *bar_rows = &bars[0], *bar_halo = &bars[1], *bar_chunk = &bars[2 + w];
        const int r0 = m.round_start[t * NW + w], nr = m.round_start[t * NW + w + 1] - r0;
        const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0, hr = m.halo_round[t * NW + w];
        constexpr unsigned kRow = T * (unsigned)sizeof(R);
        if (tid == 0) {
            mbar_init(bar_rows, 1); mbar_init(bar_halo, T);
            for (int i = 0; i < NW; i++) mbar_init(&bars[2 + i], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar_rows, 25u * kRow);
            for (int k = 0; k < 5; k++) bulk_g2s(qg + k * TS, Q + (long)k * m.sN + c0, kRow, bar_rows);
            for (int k = 0; k < 15; k++) bulk_g2s(qg + (5 + k) * TS, G + (long)k * m.sN + c0, kRow, bar_rows);
            for (int k = 0; k < 5; k++) bulk_g2s(ab + k * TS, abar + (long)k * m.sC + c0, kRow, bar_rows);
        }
        if (lane == 0 && nr > 0) {
            mbar_expect_tx(bar_chunk, (unsigned)Ch::kBytes);
            bulk_g2s(chunk, m.chunks + (long)r0 * Ch::kScalars, (unsigned)Ch::kBytes, bar_chunk);
        }
        const R vown = m.vol[c0 + tid];             // (rows past the last cell: allocated slack)
        {
            constexpr int kIter = (TS - T + T - 1) / T;
            int hc[kIter];
            #pragma unroll
            for (int i = 0; i < kIter; i++) { const int h = tid + i * T; hc[i] = h < nh ? m.halo_cell[h0 + h] : -1; }
            #pragma unroll
            for (int i = 0; i < kIter; i++) {
                const int cell = hc[i], slot = T + tid + i * T;
                if (cell < 0) continue;
                for (int k = 0; k < 5; k++) cp_async_elem<sizeof(R)>(qg + k * TS + slot, Q + (long)k * m.sN + cell);
                for (int k = 0; k < 15; k++) cp_async_elem<sizeof(R)>(qg + (5 + k) * TS + slot, G + (long)k * m.sN + cell);
                if (cell < m.nInternalCells) {
                    for (int k = 0; k < 5; k++) cp_async_elem<sizeof(R)>(ab + k * TS + slot, abar + (long)k * m.sC + cell);
                    cp_async_elem<sizeof(R)>(volh + slot - T, m.vol + cell);
                }
            }
        }
        cp_async_arrive(bar_halo);
        {   // L2 prefetch for the tile one wave later (see FluxTileBody)
            const int tn = t + kPrefetchDistance;
            if (tn < m.nTiles && w == NW - 1) {
                const long cn = (long)tn * T;
                if (lane < 5) bulk_prefetch_l2(Q + (long)lane * m.sN + cn, kRow);
                else if (lane < 20) bulk_prefetch_l2(G + (long)(lane - 5) * m.sN + cn, kRow);
                else if (lane < 25) bulk_prefetch_l2(abar + (long)(lane - 20) * m.sC + cn, kRow);
                else if (lane == 25) bulk_prefetch_l2(m.vol + cn, kRow);
            }
            if (tn < m.nTiles && w == (NW > 1 ? NW - 2 : 0)) tile_prefetch_meta<R, NW, Ch>(m, tn, tn + kPrefetchDistance, lane);
        }
        // the tile's own abar rows are consumed divided by the cell volume, by every warp of the CTA: scale them in place once
        mbar_wait(bar_rows, 0);
        const R ivh = tid < nc ? rcp(vown) : R(0);
        #pragma unroll
        for (int k = 0; k < 5; k++) ab[k * TS + tid] *= ivh;
        __syncthreads();
        if (nr == 0) return;
        Phys<R> phs = ph;
        if (SPEC == SPEC_ROE_SUTHERLAND) { phs.riemann = RIEMANN_ROE; phs.mu_law = MU_SUTHERLAND; }
        if (SPEC == SPEC_ROE_CONSTANT) { phs.riemann = RIEMANN_ROE; phs.mu_law = MU_CONSTANT; }
        R acc[20];
        for (int k = 0; k < 20; k++) acc[k] = R(0);
        for (int r = 0; r < nr; r++) {
            mbar_wait(bar_chunk, (unsigned)(r & 1));
            TileEntry e; tile_decode(Ch::words(chunk)[lane], e);
            Geom<R> gm; tile_load_geom<R, 32>(chunk, lane, gm);
            if (r == hr) mbar_wait(bar_halo, 0);
            __syncwarp();
            if (lane == 0 && r + 1 < nr) {
                fence_proxy_async();
                mbar_expect_tx(bar_chunk, (unsigned)Ch::kBytes);
                bulk_g2s(chunk, m.chunks + (long)(r0 + r + 1) * Ch::kScalars, (unsigned)Ch::kBytes, bar_chunk);
            }
            R snd[20];
            for (int k = 0; k < 20; k++) snd[k] = R(0);
            if (e.valid) face(phs, gm, e, tid, ghost_of(t, e), qg, ab, volh, acc, snd);
            if (__any_sync(0xffffffffu, e.has_src)) {
                for (int k = 0; k < 20; k++) { const R v = __shfl_sync(0xffffffffu, snd[k], e.src); if (e.has_src) acc[k] += v; }
            }
        }
        if (tid < nc) finish(c0 + tid, acc, ivh);
    }
#endif
};

// ------------------------------------------------------------------------------------------ reverse of gradCell (+ primitive + RK)
// C + F of fvm_bodies.h, one CTA per tile, one thread per cell: the rows H = Gb/V of the tile's own cells arrive by bulk
// copy, those of its halo cells (the face neighbours in other tiles) by the asynchronous gather, and the six-neighbour
// gather of 15 values each reads shared memory instead of L2 (it was bound by the latency of those loads).
// nbrSlot (mesh upload): shared-memory slot of every neighbour; bit 15 marks ghost cells, which have no gradient of their own.
enum { kGhostSlot = 0x8000 };
template <typename R> struct NbrSlotBody {
    static constexpr const char* kName = "nbr_slot";
    MeshDev<R> m; unsigned short* out;
    FVM_HD void operator()(int c) const {
        const int t = c / m.T, c0 = t * m.T, h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0;
        for (int j = 0; j < 6; j++) {
            const int nb = m.cellNbr[(long)j * m.sC + c];
            int slot = -1;
            if (nb >= c0 && nb < c0 + m.T && nb < m.nInternalCells) slot = nb - c0;
            else for (int h = 0; h < nh; h++) if (m.halo_cell[h0 + h] == nb) { slot = m.T + h; break; }
            // every face neighbour outside the tile is a halo slot of the tile (fvm_tiles.h); 0x7fff would fault loudly
            out[(long)j * m.sC + c] = (unsigned short)((slot < 0 ? 0x7fff : slot) | (nb >= m.nInternalCells ? kGhostSlot : 0));
        }
    }
};

template <typename R, int T, int TS> struct GradAdjTileBody {
    static constexpr const char* kName = "grad_adj_update";
    static constexpr int kThreads = T;
    static constexpr int kMinBlocks = sizeof(R) == 8 ? 4 : 5;      // register cap 128 (fp64) / 102 (fp32)
    static constexpr int kPrefetchDistance = 148 * kMinBlocks;
    Phys<R> ph; MeshDev<R> m;
    const R* Gb; R* Qb;
    const R* W;                // stage state the residual was evaluated at
    const R *A1, *A2, *A3;     // adjoints of later stage outputs (NULL when coefficient is 0)
    R c1, c2, c3;
    R objT;                    // obja for OBJ_CELL_TV / OBJ_CELL_T on the objective stage, else 0
    int objVol;                // 1: the cell objective is volume-weighted (OBJ_CELL_TV)
    R* Aout;                   // [5][sC]
    R* Sb; R s1, s2, s3;       // source gradient accumulation (only when Sb != NULL): Sb += s1*A1 + s2*A2 + s3*A3
#if defined(__CUDACC__)
    static constexpr size_t kBarOff = ((size_t)15 * TS * sizeof(R) + 15) / 16 * 16;
    static size_t smem_bytes() { return kBarOff + 16; }
#endif
    // per-cell inputs that do not depend on the staged rows: loaded BEFORE the CTA waits for its shared-memory rows, so that
    // the two memory round trips overlap (cell-face metrics, neighbour slots, the cell's own Qb row)
    struct Pre { R cfm[24]; int slots[6]; unsigned ghosts; Prim<R> acc; };
    FVM_HD void preload(int c, Pre& p) const {
        const int sC = m.sC;
        const R* FVM_RESTRICT cfm = m.cfm;
        p.ghosts = 0;
        for (int j = 0; j < 6; j++) {
            const int s = m.nbrSlot[(long)j * sC + c];
            p.slots[j] = s & 0x7fff;
            if (s & kGhostSlot) p.ghosts |= 1u << j;
        }
        for (int k = 0; k < 24; k++) p.cfm[k] = cfm[(long)k * sC + c];
        load_prim(Qb, m.sN, c, p.acc);
    }
    // cell c of tile t (slot lo); gbs [15][TS]: H rows of the tile's slots (ghost slots never read)
    FVM_HD void cell(int t, int c, int lo, const R* gbs, const Pre& p) const {
        const int sC = m.sC, sN = m.sN;
        R H[15];
        for (int k = 0; k < 15; k++) H[k] = gbs[k * TS + lo];
        Prim<R> acc = p.acc;
        for (int j = 0; j < 6; j++) {
            const R* SN = p.cfm + 4 * j;
            const R a = p.cfm[4 * j + 3];
            R D[15];
            if ((p.ghosts >> j) & 1u) { for (int k = 0; k < 15; k++) D[k] = H[k]; }
            else { const R* nb = gbs + p.slots[j]; for (int k = 0; k < 15; k++) D[k] = H[k] - nb[k * TS]; }
            for (int i = 0; i < 3; i++) acc.U[i] += a * (SN[0] * D[3 * i] + SN[1] * D[3 * i + 1] + SN[2] * D[3 * i + 2]);
            acc.T += a * (SN[0] * D[9] + SN[1] * D[10] + SN[2] * D[11]);
            acc.p += a * (SN[0] * D[12] + SN[1] * D[13] + SN[2] * D[14]);
        }
        if (p.ghosts) {
            for (int j = 0; j < 6; j++) {
                if (!((p.ghosts >> j) & 1u)) continue;
                const R* SN = p.cfm + 4 * j;
                const R wp = R(1) - p.cfm[4 * j + 3];
                Prim<R> gq;
                for (int i = 0; i < 3; i++) gq.U[i] = wp * (SN[0] * H[3 * i] + SN[1] * H[3 * i + 1] + SN[2] * H[3 * i + 2]);
                gq.T = wp * (SN[0] * H[9] + SN[1] * H[10] + SN[2] * H[11]);
                gq.p = wp * (SN[0] * H[12] + SN[1] * H[13] + SN[2] * H[14]);
                add_prim(Qb, sN, m.halo_cell[m.halo_start[t] + p.slots[j] - T], gq);      // the ghost row of that face (exclusive writer)
            }
        }
        if (objT != R(0)) acc.T += objVol ? objT * m.vol[c] : objT;
        R rhoU[3] = {W[sC + c], W[2 * sC + c], W[3 * sC + c]};
        R out[5] = {0, 0, 0, 0, 0};
        primitive_vjp(ph, W[c], rhoU, W[4 * sC + c], acc, out[0], out + 1, out[4]);
        for (int k = 0; k < 5; k++) {
            R v = out[k];
            R x1 = A1 ? A1[k * sC + c] : R(0), x2 = A2 ? A2[k * sC + c] : R(0), x3 = A3 ? A3[k * sC + c] : R(0);
            if (A1) v += c1 * x1;
            if (A2) v += c2 * x2;
            if (A3) v += c3 * x3;
            Aout[k * sC + c] = v;
            if (Sb) Sb[k * sC + c] += s1 * x1 + s2 * x2 + s3 * x3;
        }
    }
#if !defined(__CUDACC__)
    void host_tile(int t) const {
        const int c0 = t * T, nc = (m.nInternalCells - c0 < T) ? m.nInternalCells - c0 : T;
        const R nan = std::numeric_limits<R>::quiet_NaN();
        std::vector<R> gbs((size_t)15 * TS, nan);
        for (int l = 0; l < nc; l++) for (int k = 0; k < 15; k++) gbs[(size_t)k * TS + l] = Gb[(long)k * m.sN + c0 + l];
        for (int h = m.halo_start[t]; h < m.halo_start[t + 1]; h++) {
            const int cellh = m.halo_cell[h];
            if (cellh < m.nInternalCells) for (int k = 0; k < 15; k++) gbs[(size_t)k * TS + T + h - m.halo_start[t]] = Gb[(long)k * m.sN + cellh];
        }
        for (int l = 0; l < nc; l++) { Pre p; preload(c0 + l, p); cell(t, c0 + l, l, gbs.data(), p); }
    }
#else
    __device__ __forceinline__ void device_tile(int t, unsigned char* smem) const {
        const int c0 = t * T, nc = min(T, m.nInternalCells - c0);
        const int tid = threadIdx.x;
        R* gbs = reinterpret_cast<R*>(smem);
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kBarOff);
        unsigned long long *bar_rows = &bars[0], *bar_halo = &bars[1];
        const int h0 = m.halo_start[t], nh = m.halo_start[t + 1] - h0;
        constexpr unsigned kRow = T * (unsigned)sizeof(R);
        if (tid == 0) { mbar_init(bar_rows, 1); mbar_init(bar_halo, T); mbar_fence_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar_rows, 15u * kRow);
            for (int k = 0; k < 15; k++) bulk_g2s(gbs + k * TS, Gb + (long)k * m.sN + c0, kRow, bar_rows);
        }
        constexpr int kIter = (TS - T + T - 1) / T;
        int hc[kIter];
        #pragma unroll
        for (int i = 0; i < kIter; i++) { const int h = tid + i * T; hc[i] = h < nh ? m.halo_cell[h0 + h] : -1; }
        // the cell's own inputs do not depend on anything: issue their loads BEFORE the gather below stalls on the halo cell list,
        // and pull the lines the tail of cell() reads (stage state, later-stage adjoints, gradient accumulator) into L2 meanwhile
        Pre p;
        if (tid < nc) preload(c0 + tid, p);
        if ((tid & 15) == 0 && tid < nc) {
            const long c = c0 + tid;
            for (int k = 0; k < 5; k++) {
                prefetch_l2(W + (long)k * m.sC + c);
                if (A1) prefetch_l2(A1 + (long)k * m.sC + c);
                if (A2) prefetch_l2(A2 + (long)k * m.sC + c);
                if (A3) prefetch_l2(A3 + (long)k * m.sC + c);
                if (Sb) prefetch_l2(Sb + (long)k * m.sC + c);
            }
        }
        #pragma unroll
        for (int i = 0; i < kIter; i++) {
            const int cellh = hc[i], slot = T + tid + i * T;
            if (cellh < 0 || cellh >= m.nInternalCells) continue;
            for (int k = 0; k < 15; k++) cp_async_elem<sizeof(R)>(gbs + k * TS + slot, Gb + (long)k * m.sN + cellh);
        }
        cp_async_arrive(bar_halo);
        {   // L2 prefetch of the rows of the tile that runs in this CTA slot one wave later
            const int tn = t + kPrefetchDistance;
            if (tn < m.nTiles && tid < 15) bulk_prefetch_l2(Gb + (long)tid * m.sN + (long)tn * T, kRow);
            if (tn < m.nTiles && tid >= 32 && tid < 39) { const int* p = m.halo_cell + m.halo_start[tn] + 32 * (tid - 32); if (p < m.halo_cell + m.halo_start[tn + 1]) prefetch_l2(p); }
        }
        mbar_wait(bar_rows, 0);
        mbar_wait(bar_halo, 0);
        if (tid < nc) cell(t, c0 + tid, tid, gbs, p);
    }
#endif
};

}  // namespace fvm
