// Host-side, one-off mesh preprocessing for the tiled flux kernels (runs inside adfvm_set_mesh).
//
// The reference scatters face fluxes into cells with one atomicAdd per scalar (adpy/adpy/tensor.py:393-394) and
// visits faces in file order. Here the internal cells are regrouped into spatially compact TILES of `T`
// consecutive (renumbered) cells, each made of T/32 compact SUB-TILES of 32 consecutive cells. One CTA owns one
// tile and stages the tile's cell rows (+ halo) in shared memory once; one WARP owns one sub-tile: it evaluates
// every face touching its 32 cells (faces cut by a sub-tile boundary are evaluated by both sides, each side adds
// only its own share) and sums the contributions into shared-memory accumulators that no other warp touches, in a
// FIXED order given by a face colouring: faces of one colour never share a cell of the sub-tile, colours are
// applied one after the other. No float atomics, no block-wide barriers in the face loop, bitwise reproducible,
// and 4.0 flux evaluations per hex cell (4x4x2 sub-tiles) instead of the 6 of a cell-centred gather.
//
//   * tiles: recursive coordinate bisection of the cell centres (recovered up to a translation by walking the
//     internal faces and adding deltas*deltasUnit = N-P, reference adFVM/cpp/cmesh.cpp:184-193), always splitting
//     at a multiple of T (then of 32 inside a tile) so that every tile / sub-tile but the last is full;
//   * internal faces are renumbered in the order the sub-tiles first list them (coalesced metric loads);
//     boundary faces and ghost cells keep the reference's numbering (ghost(f) = C + f - Fi, adFVM/mesh.py:244),
//     so patch ranges and the halo layout are untouched.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace fvm {

// packed entry word: bits 0-9 owner's slot, 10-19 neighbour's slot, 20-24 colour, 25-26 face kind (FaceKind of
// fvm_math.h), 27 valid, 28 owner is a cell of this sub-tile (scatter to it), 29 neighbour is a cell of this
// sub-tile, 30 neighbour is a ghost cell (boundary face). Slots [0,T) are the tile's own cells, slots
// [T, T+nHalo) the tile's halo: cells of other tiles and ghost cells touched by the tile's faces.
enum { kRound = 32 };   // entries per warp round
static inline uint32_t tile_pack(int lo, int ln, int colour, int kind, int valid, int so = 0, int sn = 0, int ghost = 0) {
    return (uint32_t)lo | ((uint32_t)ln << 10) | ((uint32_t)colour << 20) | ((uint32_t)kind << 25) | ((uint32_t)valid << 27) |
           ((uint32_t)so << 28) | ((uint32_t)sn << 29) | ((uint32_t)ghost << 30);
}

struct TilePlan {
    int T = 0, nTiles = 0, NW = 0, maxColours = 0;   // NW = T/32 sub-tiles (warps) per tile
    long nEntries = 0;
    std::vector<int> cell_new2old, cell_old2new;     // internal cells
    std::vector<int> face_new2old, face_old2new;     // all faces (identity for boundary faces)
    std::vector<int> round_start;                    // [nTiles*NW+1] first round of each sub-tile; round r covers entry slots [32r, 32r+32)
    std::vector<int> ent_face;                       // NEW face index of each entry slot, -1 for padding
    std::vector<uint32_t> ent_loc;                   // tile_pack(...) of each entry slot
    std::vector<int> halo_start;                     // [nTiles+1] offsets into halo_cell
    std::vector<int> halo_cell;                      // NEW cell index (ghost cells: >= C) of each halo slot
    int maxHalo = 0;
    std::vector<int> halo_round;                     // [nTiles*NW] first round of the sub-tile (relative) that reads a halo slot
    double evals_per_cell() const { return cell_new2old.empty() ? 0. : (double)nEntries / (double)cell_new2old.size(); }
};

namespace detail {

// cell centres up to a translation per connected component: x[nbr] = x[owner] + delta * deltaUnit
template <typename R>
void integrate_centres(int C, int Fi, const int* owner, const int* neigh, const int* cellFaces, const R* deltas,
                       const R* deltasUnit, std::vector<float>& pos) {
    pos.assign((size_t)3 * C, 0.f);
    std::vector<char> seen(C, 0);
    std::vector<int> queue; queue.reserve(C);
    std::vector<double> p((size_t)3 * C, 0.);
    double shift = 0.;
    for (int seed = 0; seed < C; seed++) {
        if (seen[seed]) continue;
        seen[seed] = 1; p[3 * (size_t)seed] = shift; queue.clear(); queue.push_back(seed);
        double xmax = shift;
        for (size_t h = 0; h < queue.size(); h++) {
            const int c = queue[h];
            for (int j = 0; j < 6; j++) {
                const int f = cellFaces[(size_t)c * 6 + j];
                if (f >= Fi) continue;
                const int o = owner[f], n = neigh[f];
                const int other = (o == c) ? n : o;
                if (seen[other]) continue;
                const double sg = (o == c) ? 1. : -1.;
                for (int k = 0; k < 3; k++)
                    p[3 * (size_t)other + k] = p[3 * (size_t)c + k] + sg * (double)deltas[f] * (double)deltasUnit[(size_t)f * 3 + k];
                xmax = std::max(xmax, p[3 * (size_t)other]);
                seen[other] = 1; queue.push_back(other);
            }
        }
        shift = xmax + 1.;     // next component goes beside this one
    }
    for (size_t i = 0; i < p.size(); i++) pos[i] = (float)p[i];
}

// recursive coordinate bisection of idx[lo,hi) into leaves of T cells (splits at multiples of T)
inline void rcb(std::vector<int>& idx, const std::vector<float>& pos, long lo, long hi, int T) {
    struct Range { long lo, hi; };
    std::vector<Range> stack; stack.push_back({lo, hi});
    while (!stack.empty()) {
        Range r = stack.back(); stack.pop_back();
        const long n = r.hi - r.lo;
        if (n <= T) continue;
        float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
        for (long i = r.lo; i < r.hi; i++)
            for (int k = 0; k < 3; k++) { const float v = pos[3 * (size_t)idx[i] + k]; mn[k] = std::min(mn[k], v); mx[k] = std::max(mx[k], v); }
        int dim = 0;
        for (int k = 1; k < 3; k++) if (mx[k] - mn[k] > (mx[dim] - mn[dim]) * 1.0001f) dim = k;
        const long tiles = (n + T - 1) / T;
        const long left = (tiles / 2) * T;
        // order by (coordinate, index): deterministic under ties
        std::nth_element(idx.begin() + r.lo, idx.begin() + r.lo + left, idx.begin() + r.hi, [&](int a, int b) {
            const float pa = pos[3 * (size_t)a + dim], pb = pos[3 * (size_t)b + dim];
            return pa != pb ? pa < pb : a < b; });
        stack.push_back({r.lo + left, r.hi});
        stack.push_back({r.lo, r.lo + left});
    }
}


// Lane assignment inside one round of 32 entries (8-byte scalars). A warp-wide LDS.64 is served half-warp by
// half-warp; lanes of one half that read DIFFERENT cell slots with the same slot%16 (the same pair of 4-byte banks)
// serialise. The accumulation order is fixed by the colours, not by the lanes, so the lanes of a round may be
// permuted freely: pick the split into two halves that minimises, for the owner-side and the neighbour-side loads,
// the largest number of distinct slots per bank pair (greedy start + pairwise-swap descent; identical rounds of a
// structured mesh are memoised).
struct RoundBalancer {
    std::unordered_map<uint64_t, std::vector<std::pair<std::vector<uint32_t>, std::vector<uint8_t>>>> memo;
    // per half h, side s (0 owner slot, 1 neighbour slot): refs[h][s][slot] entries reading the slot, distinct[h][s][res]
    uint8_t refs[2][2][1024]; uint8_t distinct[2][2][16];
    static int slot_of(uint32_t w, int side) { return side ? (int)((w >> 10) & 0x3FFu) : (int)(w & 0x3FFu); }
    void add(int h, uint32_t w, int d) {
        if (!((w >> 27) & 1u)) return;                      // padding entries do not load
        for (int s = 0; s < 2; s++) {
            const int sl = slot_of(w, s);
            if (d > 0) { if (refs[h][s][sl]++ == 0) distinct[h][s][sl & 15]++; }
            else { if (--refs[h][s][sl] == 0) distinct[h][s][sl & 15]--; }
        }
    }
    int cost() const {
        int c = 0;
        for (int h = 0; h < 2; h++) for (int s = 0; s < 2; s++) {
            int mx = 0, sum = 0;
            for (int r = 0; r < 16; r++) { const int d = distinct[h][s][r]; mx = std::max(mx, d); sum += d * d; }
            c += 64 * mx + sum;                             // wavefronts first, spread as the tie-break
        }
        return c;
    }
    void balance(uint32_t* loc, int* face) {
        uint64_t key = 1469598103934665603ull;
        for (int i = 0; i < kRound; i++) { key ^= loc[i] & 0x080FFFFFu; key *= 1099511628211ull; }
        std::vector<uint32_t> sig(loc, loc + kRound);
        for (auto& x : sig) x &= 0x080FFFFFu;
        std::vector<uint8_t> perm;
        auto& bucket = memo[key];
        for (auto& kv : bucket) if (kv.first == sig) { perm = kv.second; break; }
        if (perm.empty()) {
            std::memset(refs, 0, sizeof(refs)); std::memset(distinct, 0, sizeof(distinct));
            int half[kRound], n[2] = {0, 0};
            for (int i = 0; i < kRound; i++) {              // greedy: the half where the entry adds less
                int best = -1, bc = 0;
                for (int h = 0; h < 2; h++) {
                    if (n[h] >= kRound / 2) continue;
                    add(h, loc[i], +1); const int c = cost() + n[h]; add(h, loc[i], -1);
                    if (best < 0 || c < bc) { best = h; bc = c; }
                }
                half[i] = best; n[best]++; add(best, loc[i], +1);
            }
            int cur = cost();
            for (int sweep = 0; sweep < 4; sweep++) {
                bool improved = false;
                for (int i = 0; i < kRound; i++) for (int j = i + 1; j < kRound; j++) {
                    if (half[i] == half[j]) continue;
                    add(half[i], loc[i], -1); add(half[j], loc[j], -1); add(half[j], loc[i], +1); add(half[i], loc[j], +1);
                    const int c = cost();
                    if (c < cur) { cur = c; std::swap(half[i], half[j]); improved = true; }
                    else { add(half[j], loc[i], -1); add(half[i], loc[j], -1); add(half[i], loc[i], +1); add(half[j], loc[j], +1); }
                }
                if (!improved) break;
            }
            perm.resize(kRound);
            int pos[2] = {0, kRound / 2};
            for (int i = 0; i < kRound; i++) perm[pos[half[i]]++] = (uint8_t)i;    // new lane -> old lane
            bucket.push_back({sig, perm});
        }
        uint32_t l2[kRound]; int f2[kRound];
        for (int i = 0; i < kRound; i++) { l2[i] = loc[perm[i]]; f2[i] = face[perm[i]]; }
        for (int i = 0; i < kRound; i++) { loc[i] = l2[i]; face[i] = f2[i]; }
    }
};

}  // namespace detail

// owner/neigh/cellFaces: the reference's arrays (old numbering). T: multiple of 32, <= 512.
// bkind[f - Fi]: FaceKind of each boundary face.
template <typename R>
TilePlan build_tile_plan(int C, int Fi, int F, const int* owner, const int* neigh, const int* cellFaces,
                         const R* deltas, const R* deltasUnit, const unsigned char* bkind, int T) {
    if (T <= 0 || T > 512 || T % kRound) throw std::runtime_error("tile size out of range");
    TilePlan P; P.T = T; P.NW = T / kRound;
    P.nTiles = (C + T - 1) / T;
    const int NW = P.NW;
    // ---- cell order: tiles of T cells, inside a tile sub-tiles of 32 cells, inside a sub-tile ascending old ids
    std::vector<float> pos;
    detail::integrate_centres(C, Fi, owner, neigh, cellFaces, deltas, deltasUnit, pos);
    P.cell_new2old.resize(C);
    std::iota(P.cell_new2old.begin(), P.cell_new2old.end(), 0);
    detail::rcb(P.cell_new2old, pos, 0, C, T);
    for (int t = 0; t < P.nTiles; t++) detail::rcb(P.cell_new2old, pos, (long)t * T, std::min<long>((long)(t + 1) * T, C), kRound);
    for (long b = 0; b < C; b += kRound)
        std::sort(P.cell_new2old.begin() + b, P.cell_new2old.begin() + std::min<long>(b + kRound, C));
    P.cell_old2new.assign(C, -1);
    for (int i = 0; i < C; i++) P.cell_old2new[P.cell_new2old[i]] = i;
    // ---- per sub-tile: faces touching it, coloured
    P.round_start.assign((size_t)P.nTiles * NW + 1, 0);
    P.halo_round.assign((size_t)P.nTiles * NW, 0);
    P.face_old2new.assign(F, -1);
    P.face_new2old.assign(F, -1);
    for (int f = Fi; f < F; f++) { P.face_old2new[f] = f; P.face_new2old[f] = f; }
    int nextFace = 0;
    std::vector<int> faces, colour, group, order;
    std::vector<uint32_t> used(kRound);
    static thread_local detail::RoundBalancer balancer_store; detail::RoundBalancer& balancer = balancer_store; balancer.memo.clear();
    const int N = C + (F - Fi);
    std::vector<int> slot_of(N, -1), slot_tile(N, -1);
    P.halo_start.assign(P.nTiles + 1, 0);
    for (int t = 0; t < P.nTiles; t++) {
        const int c0 = t * T, c1 = std::min(C, c0 + T);
        int nHalo = 0;
        auto slot = [&](int cell) {           // cell: NEW index (ghosts >= C)
            if (cell >= c0 && cell < c1) return cell - c0;
            if (slot_tile[cell] != t) { slot_tile[cell] = t; slot_of[cell] = T + nHalo++; P.halo_cell.push_back(cell); }
            return slot_of[cell];
        };
        for (int w = 0; w < NW; w++) {
            const int s0 = c0 + w * kRound, s1 = std::min(c1, s0 + kRound);
            faces.clear();
            for (int c = s0; c < s1; c++) {
                const int oc = P.cell_new2old[c];
                for (int j = 0; j < 6; j++) {
                    const int f = cellFaces[(size_t)oc * 6 + j];
                    if (f < 0 || f >= F) throw std::runtime_error("cellFaces out of range");
                    if (f >= Fi) { faces.push_back(f); continue; }
                    // internal face: list once per sub-tile (when reached from its lower cell of the sub-tile)
                    const int a = P.cell_old2new[owner[f]], b = P.cell_old2new[neigh[f]];
                    const bool ain = a >= s0 && a < s1, bin = b >= s0 && b < s1;
                    if (ain && bin) { if (c == std::min(a, b)) faces.push_back(f); }
                    else faces.push_back(f);
                }
            }
            // greedy colouring: no two faces of one colour share a cell of the sub-tile. Faces inside the sub-tile are
            // coloured first, then faces whose other cell is elsewhere in the tile, then faces that need a halo slot (cut
            // by the tile boundary, or boundary faces): sorted by colour, the rounds that only need the tile's own rows
            // come first and overlap the halo gather.
            colour.assign(faces.size(), 0); group.assign(faces.size(), 0);
            for (size_t i = 0; i < faces.size(); i++) {
                const int f = faces[i];
                const int a = P.cell_old2new[owner[f]];
                const int b = f < Fi ? P.cell_old2new[neigh[f]] : -1;
                const bool ain = a >= s0 && a < s1, bin = b >= s0 && b < s1;
                const bool atile = a >= c0 && a < c1, btile = b >= c0 && b < c1;
                group[i] = (ain && bin) ? 0 : ((atile && btile) ? 1 : 2);
            }
            int ncol = 0;
            for (int grp = 0; grp < 3; grp++) {
                std::fill(used.begin(), used.end(), 0u);
                const int base = ncol;
                for (size_t i = 0; i < faces.size(); i++) {
                    if (group[i] != grp) continue;
                    const int f = faces[i];
                    const int a = P.cell_old2new[owner[f]];
                    const int b = f < Fi ? P.cell_old2new[neigh[f]] : -1;
                    const int la = (a >= s0 && a < s1) ? a - s0 : -1, lb = (b >= s0 && b < s1) ? b - s0 : -1;
                    uint32_t mk = (la >= 0 ? used[la] : 0u) | (lb >= 0 ? used[lb] : 0u);
                    int col = 0;
                    while (mk & (1u << col)) col++;
                    if (base + col >= 32) throw std::runtime_error("face colouring needs more than 32 colours");
                    colour[i] = base + col; ncol = std::max(ncol, base + col + 1);
                    if (la >= 0) used[la] |= 1u << col;
                    if (lb >= 0) used[lb] |= 1u << col;
                }
            }
            P.maxColours = std::max(P.maxColours, ncol);
            order.resize(faces.size());
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return colour[x] < colour[y]; });
            int firstHaloEntry = -1, nEnt = 0;
            for (int i : order) {
                const int f = faces[i];
                if (f < Fi && P.face_old2new[f] < 0) { P.face_old2new[f] = nextFace; P.face_new2old[nextFace] = f; nextFace++; }
                const int a = P.cell_old2new[owner[f]];
                const int b = f < Fi ? P.cell_old2new[neigh[f]] : neigh[f];
                if (b < 0 || b >= N) throw std::runtime_error("neighbour out of range");
                const int la = slot(a), lb = slot(b);
                if (la >= 1024 || lb >= 1024) throw std::runtime_error("tile halo too large");
                if ((la >= T || lb >= T) && firstHaloEntry < 0) firstHaloEntry = nEnt;
                P.ent_face.push_back(P.face_old2new[f]);
                P.ent_loc.push_back(tile_pack(la, lb, colour[i], f < Fi ? 0 : (int)bkind[f - Fi], 1,
                                              a >= s0 && a < s1, b >= s0 && b < s1, b >= C));
                P.nEntries++; nEnt++;
            }
            P.halo_round[(size_t)t * NW + w] = firstHaloEntry < 0 ? (nEnt + kRound - 1) / kRound : firstHaloEntry / kRound;
            // pad the last round (padding repeats the last colour: the kernels scan the colours present in a round)
            while (P.ent_face.size() % (size_t)kRound) { P.ent_face.push_back(-1); P.ent_loc.push_back(tile_pack(0, 0, ncol ? ncol - 1 : 0, 0, 0)); }
            if (sizeof(R) == 8)
                for (size_t e = (size_t)P.round_start[(size_t)t * NW + w] * kRound; e < P.ent_face.size(); e += kRound)
                    balancer.balance(&P.ent_loc[e], &P.ent_face[e]);
            P.round_start[(size_t)t * NW + w + 1] = (int)(P.ent_face.size() / (size_t)kRound);
        }
        P.maxHalo = std::max(P.maxHalo, nHalo);
        P.halo_start[t + 1] = (int)P.halo_cell.size();
    }
    if (nextFace != Fi) throw std::runtime_error("internal face not reachable from any cell (broken cellFaces)");
    return P;
}

}  // namespace fvm
