// Host-side, one-off mesh preprocessing for the tiled flux kernels (runs inside adfvm_set_mesh).
//
// The reference scatters face fluxes into cells with one atomicAdd per scalar (adpy/adpy/tensor.py:393-394) and
// visits faces in file order. Here the internal cells are regrouped into spatially compact TILES of `T`
// consecutive (renumbered) cells, each made of T/32 compact SUB-TILES of 32 consecutive cells. One CTA owns one
// tile and stages the tile's cell rows (+ halo) in shared memory once; one WARP owns one sub-tile and LANE l OWNS
// CELL l of it: the lane keeps that cell's accumulators in registers for the whole kernel.
//   * every face touching the sub-tile is one ENTRY (faces cut by a sub-tile boundary are evaluated by both sides);
//     an entry has a HOME cell inside the sub-tile; the lane of the home cell evaluates it, always as the "owner"
//     side: when the home cell is the face's neighbour the entry's metrics are stored FLIPPED (normal and delta
//     reversed, the two sides' reconstruction weights swapped), which turns the neighbour's share -F*A/V into the
//     owner's share +F'*A/V of the flipped face (all fluxes on the path are antisymmetric under the flip);
//   * if the other cell of an entry is in the sub-tile too, its share travels by a warp shuffle: in each round a lane
//     evaluates at most one entry and receives at most one share;
//   * the schedule is an orientation of the sub-tile's inner faces (which end is home) that balances the number of
//     entries per lane (augmenting paths), followed by a bipartite edge colouring (home lane x receiving lane) with
//     as many colours = ROUNDS as the largest degree (Koenig). A 4x4x2 hex sub-tile needs exactly 4 full rounds:
//     4.0 flux evaluations per cell instead of the 6 of a cell-centred gather, no idle lanes.
// No float atomics, no shared-memory accumulators, no barriers in the face loop; the summation order of a cell
// (round by round, own share then received share) is fixed: results are bitwise reproducible.
//
//   * tiles: recursive coordinate bisection of the cell centres (recovered up to a translation by walking the
//     internal faces and adding deltas*deltasUnit = N-P, reference adFVM/cpp/cmesh.cpp:184-193), always splitting
//     at a multiple of T (then of 32 inside a tile) so that every tile / sub-tile but the last is full;
//   * internal faces are renumbered in the order the sub-tiles first list them (coalesced metric loads);
//     boundary faces and ghost cells keep the reference's numbering (ghost(f) = C + f - Fi, adFVM/mesh.py:244),
//     so patch ranges and the halo layout are untouched.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <unordered_map>
#include <thread>
#include <atomic>
#include <exception>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace fvm {

// packed entry word of lane l in a round: bits 0-4 lane whose share this lane receives in this round, 5 "receives",
// 10-19 slot of the entry's OTHER cell (the home cell's slot is implicit: 32*warp + lane), 25-26 face kind (FaceKind
// of fvm_math.h), 27 valid, 29 the other cell is a cell of this sub-tile (its share is sent by shuffle), 30 the other
// cell is a ghost cell (boundary face), 31 metrics stored flipped (only read by FillChunksBody). Slots [0,T) are the
// tile's own cells, slots [T, T+nHalo) the tile's halo: cells of other tiles and ghost cells touched by the tile.
enum { kRound = 32 };   // lanes per round
static inline uint32_t tile_pack(int ln, int kind, int valid, int sn, int ghost, int flip, int has_src, int src) {
    return (uint32_t)src | ((uint32_t)has_src << 5) | ((uint32_t)ln << 10) | ((uint32_t)kind << 25) | ((uint32_t)valid << 27) |
           ((uint32_t)sn << 29) | ((uint32_t)ghost << 30) | ((uint32_t)flip << 31);
}

struct TilePlan {
    int nEarly = 0;                                  // tiles [0,nEarly) touch no processor-patch data (see build_tile_plan)
    int T = 0, nTiles = 0, NW = 0, maxColours = 0;   // NW = T/32 sub-tiles (warps) per tile; maxColours = most rounds of a sub-tile
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

// host threads for the one-off plan (ADFVM_PLAN_THREADS overrides; 1 = the sequential code path, which gives the identical plan)
inline int plan_threads() {
    if (const char* e = std::getenv("ADFVM_PLAN_THREADS")) return std::max(1, std::atoi(e));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
}
// fn(lo, hi) over [0, n) split into contiguous chunks, one per thread; exceptions are rethrown in the caller
template <class F> inline void parallel_for(long n, long grain, F&& fn) {
    int nt = plan_threads();
    if (n < 2 * grain) nt = 1;
    nt = (int)std::min<long>(nt, std::max<long>(1, n / std::max<long>(1, grain)));
    if (nt <= 1) { fn(0L, n); return; }
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err(nt);
    const long chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const long lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([&, t, lo, hi] { try { fn(lo, hi); } catch (...) { err[t] = std::current_exception(); } });
    }
    for (auto& x : th) x.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
}

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

// Axis-aligned lattice detection: when the cell centres of the mesh sit on planes x = x_i, y = y_j, z = z_k (uniform or
// graded blocks, boxes with holes), lat[3c+d] = index of cell c's plane along d. Returns false for anything else
// (warped or unstructured meshes), which then use plain coordinate bisection.
inline bool detect_lattice(const std::vector<float>& pos, int C, std::vector<int>& lat) {
    if (C < 64) return false;
    lat.assign((size_t)3 * C, 0);
    long cap = 2000000;
    if (const char* e = std::getenv("ADFVM_LATTICE_SAMPLE")) cap = std::max(64L, std::atol(e));     // tests exercise the sampling path on small meshes
    const long nsample = std::min<long>(C, cap);
    std::atomic<bool> off(false);
    // the three directions are independent: one thread each, the per-cell pass of each split further when threads are left
    const int inner = std::max(1, plan_threads() / 3);
    auto one_dim = [&](int d) {
        std::vector<float> sample, planes;
        sample.reserve(nsample);
        // scattered sample (a regular stride would alias with the lattice and miss whole planes); a missed plane is
        // caught by the verification pass below, which then rejects the lattice
        for (long i = 0; i < nsample; i++) sample.push_back(pos[3 * (size_t)(nsample == C ? i : (long)(((unsigned long long)i * 2654435761ull) % (unsigned long long)C)) + d]);
        std::sort(sample.begin(), sample.end());
        const float span = sample.back() - sample.front();
        const float tol = std::max(span * 1e-5f, 1e-30f);
        for (size_t i = 0; i < sample.size();) {              // clusters of values closer than tol = one plane
            size_t k = i; double sum = 0;
            while (k < sample.size() && sample[k] - sample[i] <= tol) sum += sample[k++];
            planes.push_back((float)(sum / (double)(k - i)));
            i = k;
            if (planes.size() > 8192) { off = true; return; }
        }
        if ((double)planes.size() * planes.size() * planes.size() > 64.0 * C && planes.size() > 64) { off = true; return; }   // not a lattice
        auto cells = [&](long lo, long hi) {
            for (long c = lo; c < hi && !off; c++) {
                const float x = pos[3 * (size_t)c + d];
                size_t k = std::lower_bound(planes.begin(), planes.end(), x) - planes.begin();
                if (k == planes.size() || (k > 0 && x - planes[k - 1] < planes[k] - x)) k--;
                if (std::fabs(x - planes[k]) > 2 * tol) { off = true; return; }
                lat[3 * (size_t)c + d] = (int)k;
            }
        };
        if (inner <= 1 || C < (1 << 18)) { cells(0, C); return; }
        std::vector<std::thread> th;
        const long chunk = ((long)C + inner - 1) / inner;
        for (int t = 0; t < inner; t++) { const long lo = t * chunk, hi = std::min<long>(C, lo + chunk); if (lo < hi) th.emplace_back(cells, lo, hi); }
        for (auto& x : th) x.join();
    };
    if (plan_threads() >= 3) {
        std::thread t1(one_dim, 1), t2(one_dim, 2);
        one_dim(0);
        t1.join(); t2.join();
    } else {
        for (int d = 0; d < 3 && !off; d++) one_dim(d);
    }
    return !off;
}

// Recursive bisection of idx[lo,hi) into leaves of T cells, always splitting at a multiple of T so that every leaf but
// the last is full. With lattice indices (lat != NULL) the cut is a lattice plane whose index, counted from the
// range's first plane, is a multiple of 8 (else 4, 2, 1) as close to the middle as possible: blocks whose extents are
// multiples of 4 / 8 then decompose into exact 4x4x8 tiles and 4x4x2 sub-tiles whatever the block size (a plain
// halving of 368 = 16*23 planes ends in 23-plane slabs and ragged tiles). Ranges without such a plane, and meshes
// without a lattice, are cut by cell count at the median coordinate (nth_element).
// one cut of idx[r.lo, r.hi): false when the range is a leaf (<= T cells); else the two halves in a, b
struct RcbRange { long lo, hi; };
inline bool rcb_split(std::vector<int>& idx, const std::vector<float>& pos, const int* lat, RcbRange r, int T, std::vector<long>& hist,
                      RcbRange& a, RcbRange& b) {
    const long n = r.hi - r.lo;
    if (n <= T) return false;
    if (lat) {
        int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {-1, -1, -1};
        for (long i = r.lo; i < r.hi; i++)
            for (int k = 0; k < 3; k++) { const int v = lat[3 * (size_t)idx[i] + k]; mn[k] = std::min(mn[k], v); mx[k] = std::max(mx[k], v); }
        int order[3] = {0, 1, 2};
        std::sort(order, order + 3, [&](int x, int y) { return (mx[x] - mn[x]) != (mx[y] - mn[y]) ? (mx[x] - mn[x]) > (mx[y] - mn[y]) : x < y; });
        for (int oi = 0; oi < 3; oi++) {
            const int dim = order[oi], ext = mx[dim] - mn[dim] + 1;
            if (ext < 2) break;
            hist.assign(ext + 1, 0);
            for (long i = r.lo; i < r.hi; i++) hist[lat[3 * (size_t)idx[i] + dim] - mn[dim] + 1]++;
            for (int p = 1; p <= ext; p++) hist[p] += hist[p - 1];           // hist[p] = cells on planes < mn+p
            int best = -1, bestClass = -1; long bestDist = 0;
            for (int p = 1; p < ext; p++) {
                if (hist[p] % T || hist[p] == 0 || hist[p] == n) continue;
                if (hist[p] * 5 < n || hist[p] * 5 > 4 * n) continue;          // keep the two sides within 1:4
                const int cls = (p % 8 == 0) ? 3 : (p % 4 == 0) ? 2 : (p % 2 == 0) ? 1 : 0;
                const long dist = std::labs(2 * hist[p] - n);
                if (cls > bestClass || (cls == bestClass && dist < bestDist)) { best = p; bestClass = cls; bestDist = dist; }
            }
            if (best < 0) continue;
            const int cut = mn[dim] + best;
            auto mid = std::stable_partition(idx.begin() + r.lo, idx.begin() + r.hi, [&](int c) { return lat[3 * (size_t)c + dim] < cut; });
            const long left = mid - (idx.begin() + r.lo);
            a = {r.lo, r.lo + left}; b = {r.lo + left, r.hi};
            return true;
        }
    }
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (long i = r.lo; i < r.hi; i++)
        for (int k = 0; k < 3; k++) { const float v = pos[3 * (size_t)idx[i] + k]; mn[k] = std::min(mn[k], v); mx[k] = std::max(mx[k], v); }
    int dim = 0;
    for (int k = 1; k < 3; k++) if (mx[k] - mn[k] > (mx[dim] - mn[dim]) * 1.0001f) dim = k;
    const long tiles = (n + T - 1) / T;
    const long left = (tiles / 2) * T;
    // order by (coordinate, index): deterministic under ties
    std::nth_element(idx.begin() + r.lo, idx.begin() + r.lo + left, idx.begin() + r.hi, [&](int x, int y) {
        const float pa = pos[3 * (size_t)x + dim], pb = pos[3 * (size_t)y + dim];
        return pa != pb ? pa < pb : x < y; });
    a = {r.lo, r.lo + left}; b = {r.lo + left, r.hi};
    return true;
}
inline void rcb(std::vector<int>& idx, const std::vector<float>& pos, const int* lat, long lo, long hi, int T) {
    std::vector<RcbRange> stack; stack.push_back({lo, hi});
    std::vector<long> hist;
    RcbRange a, b;
    while (!stack.empty()) {
        const RcbRange r = stack.back(); stack.pop_back();
        if (rcb_split(idx, pos, lat, r, T, hist, a, b)) { stack.push_back(b); stack.push_back(a); }
    }
}
// the same cuts (a range's cut depends on nothing outside it): the first levels one after the other, then the sub-ranges on threads
inline void rcb_parallel(std::vector<int>& idx, const std::vector<float>& pos, const int* lat, long lo, long hi, int T) {
    const int nt = plan_threads();
    if (nt <= 1 || hi - lo < (1L << 18)) { rcb(idx, pos, lat, lo, hi, T); return; }
    std::vector<RcbRange> work; work.push_back({lo, hi});
    std::vector<long> hist;
    while ((int)work.size() < 4 * nt) {
        std::vector<RcbRange> next; bool any = false;
        // the ranges of one level are disjoint: cut them on threads too (two at the second level, four at the third, ...)
        std::vector<RcbRange> as(work.size()), bs(work.size()); std::vector<char> ok(work.size(), 0);
        parallel_for((long)work.size(), 1, [&](long l0, long l1) {
            std::vector<long> h;
            for (long i = l0; i < l1; i++) ok[i] = rcb_split(idx, pos, lat, work[i], T, h, as[i], bs[i]) ? 1 : 0;
        });
        for (size_t i = 0; i < work.size(); i++) { if (ok[i]) { next.push_back(as[i]); next.push_back(bs[i]); any = true; } }
        if (!any) return;
        work.swap(next);            // (leaves drop out: nothing left to do for them)
    }
    parallel_for((long)work.size(), 1, [&](long l0, long l1) { for (long i = l0; i < l1; i++) rcb(idx, pos, lat, work[i].lo, work[i].hi, T); });
}

}  // namespace detail

namespace detail {

// Schedule of one sub-tile: entries -> (round, lane). home[e]/other[e]: local cell (0..31) of the entry's home /
// of its other cell when that is in the sub-tile too (else -1); both-in entries may be re-oriented (swap).
struct SubTileSchedule {
    struct Ent { int a, b; bool both; int pref; };   // a: home candidate, b: other candidate (both-in) or -1; pref: 0 early round, 1 late
    std::vector<Ent> E;
    std::vector<int> home, other, round;
    int R = 0;
    void solve() {
        const int n = (int)E.size();
        home.assign(n, -1); other.assign(n, -1); round.assign(n, -1);
        int load[kRound], inner[kRound], fixedLoad[kRound];
        for (int l = 0; l < kRound; l++) load[l] = inner[l] = fixedLoad[l] = 0;
        for (int e = 0; e < n; e++) {
            if (E[e].both) { inner[E[e].a]++; inner[E[e].b]++; }
            else fixedLoad[E[e].a]++;
        }
        int Rmin = (n + kRound - 1) / kRound;
        for (int l = 0; l < kRound; l++) Rmin = std::max(Rmin, fixedLoad[l]);
        if (Rmin < 1) Rmin = 1;
        // orientation: greedy, then augmenting paths until home loads <= R and received shares <= R
        for (int l = 0; l < kRound; l++) load[l] = fixedLoad[l];
        for (int e = 0; e < n; e++) {
            if (!E[e].both) { home[e] = E[e].a; other[e] = -1; continue; }
            const int a = E[e].a, b = E[e].b;
            if (load[a] <= load[b]) { home[e] = a; other[e] = b; load[a]++; } else { home[e] = b; other[e] = a; load[b]++; }
        }
        std::vector<std::vector<int>> inc(kRound);
        for (int e = 0; e < n; e++) if (E[e].both) { inc[E[e].a].push_back(e); inc[E[e].b].push_back(e); }
        for (R = Rmin;; R++) {
            // a cell hands over load along edges homed at it; BFS to a cell with room. Lower bound: received = inner - (load - fixed) <= R
            bool ok = true;
            for (int guard = 0; guard < 4096; guard++) {
                int u = -1; bool over = true;
                for (int l = 0; l < kRound && u < 0; l++) if (load[l] > R) { u = l; over = true; }
                for (int l = 0; l < kRound && u < 0; l++) if (inner[l] - (load[l] - fixedLoad[l]) > R) { u = l; over = false; }
                if (u < 0) break;
                // over: push one unit from u to some v with load[v] < R (and that still satisfies its own bounds);
                // under (too many received): pull one unit into u from some v with load[v]-1 still fine
                int prevE[kRound], seen[kRound]; for (int l = 0; l < kRound; l++) { prevE[l] = -1; seen[l] = 0; }
                std::vector<int> q; q.push_back(u); seen[u] = 1; int found = -1;
                for (size_t h = 0; h < q.size() && found < 0; h++) {
                    const int x = q[h];
                    for (int e : inc[x]) {
                        const bool homedAtX = home[e] == x;
                        if (over != homedAtX) continue;          // push along edges homed at x; pull along edges homed at the other end
                        const int y = (E[e].a == x) ? E[e].b : E[e].a;
                        if (seen[y]) continue;
                        seen[y] = 1; prevE[y] = e; q.push_back(y);
                        const bool room = over ? (load[y] < R) : (load[y] - 1 >= 0 && inner[y] - (load[y] - 1 - fixedLoad[y]) <= R && load[y] > fixedLoad[y]);
                        if (room) { found = y; break; }
                    }
                }
                if (found < 0) { ok = false; break; }
                // flip the path found -> ... -> u
                int y = found;
                while (y != u) {
                    const int e = prevE[y];
                    const int x = (E[e].a == y) ? E[e].b : E[e].a;
                    if (over) { home[e] = y; other[e] = x; load[y]++; load[x]--; }     // edge was homed at x: moves to y
                    else { home[e] = x; other[e] = y; load[x]++; load[y]--; }           // edge was homed at y: moves to x
                    y = x;
                }
            }
            if (!ok) continue;
            bool fine = true;
            for (int l = 0; l < kRound; l++) if (load[l] > R || inner[l] - (load[l] - fixedLoad[l]) > R) fine = false;
            if (fine) break;
        }
        // bipartite edge colouring with R colours: left = home lane, right = receiving lane
        std::vector<int> colL((size_t)kRound * R, -1), colR((size_t)kRound * R, -1);
        auto freeL = [&](int u, bool late) { if (late) { for (int c = R - 1; c >= 0; c--) if (colL[(size_t)u * R + c] < 0) return c; }
                                             else for (int c = 0; c < R; c++) if (colL[(size_t)u * R + c] < 0) return c; return -1; };
        std::vector<int> order(n); std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return E[x].pref < E[y].pref; });
        for (int e : order) {
            const int u = home[e], v = other[e];
            const bool late = E[e].pref > 0;
            if (v < 0) { const int c = freeL(u, late); if (c < 0) throw std::runtime_error("tile schedule: no free round"); colL[(size_t)u * R + c] = e; round[e] = c; continue; }
            int both = -1;
            for (int c = 0; c < R; c++) if (colL[(size_t)u * R + c] < 0 && colR[(size_t)v * R + c] < 0) { both = c; break; }
            if (both < 0) {
                int ca = -1, cb = -1;
                for (int c = 0; c < R; c++) { if (ca < 0 && colL[(size_t)u * R + c] < 0) ca = c; if (cb < 0 && colR[(size_t)v * R + c] < 0) cb = c; }
                if (ca < 0 || cb < 0) throw std::runtime_error("tile schedule: degree exceeds the number of rounds");
                // alternating ca/cb path starting at right vertex v (ca busy there): swap the colours along it
                std::vector<int> path; int x = v; bool right = true; int want = ca;
                for (;;) {
                    const int pe = right ? colR[(size_t)x * R + want] : colL[(size_t)x * R + want];
                    if (pe < 0) break;
                    path.push_back(pe);
                    x = right ? home[pe] : other[pe];
                    right = !right; want = (want == ca) ? cb : ca;
                    if (x < 0) break;                             // reached an entry without a receiving end
                }
                for (int pe : path) { colL[(size_t)home[pe] * R + round[pe]] = -1; if (other[pe] >= 0) colR[(size_t)other[pe] * R + round[pe]] = -1; }
                for (int pe : path) { round[pe] = (round[pe] == ca) ? cb : ca; }
                for (int pe : path) { colL[(size_t)home[pe] * R + round[pe]] = pe; if (other[pe] >= 0) colR[(size_t)other[pe] * R + round[pe]] = pe; }
                both = ca;
                if (colL[(size_t)u * R + both] >= 0 || colR[(size_t)v * R + both] >= 0) throw std::runtime_error("tile schedule: alternating path failed");
            }
            colL[(size_t)u * R + both] = e; colR[(size_t)v * R + both] = e; round[e] = both;
        }
    }
};

}  // namespace detail

// owner/neigh/cellFaces: the reference's arrays (old numbering). T: multiple of 32, <= 512.
// bkind[f - Fi]: FaceKind of each boundary face. Faces [nLocalFaces, F) belong to processor patches: tiles that hold,
// as own or halo cell, a cell next to such a face (hence every tile that reads or writes a remote ghost row) are LATE
// and numbered after the EARLY ones, so that early tiles / their cells can be processed while the halo is in flight.
template <typename R>
TilePlan build_tile_plan(int C, int Fi, int F, const int* owner, const int* neigh, const int* cellFaces,
                         const R* deltas, const R* deltasUnit, const unsigned char* bkind, int T, int nLocalFaces) {
    if (T <= 0 || T > 512 || T % kRound) throw std::runtime_error("tile size out of range");
    const bool verbose = std::getenv("ADFVM_TILE_TIMING") != nullptr;
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (verbose) { auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[tile plan] %-22s %.2f s\n", what, std::chrono::duration<double>(now - t_start).count()); t_start = now; } };
    TilePlan P; P.T = T; P.NW = T / kRound;
    P.nTiles = (C + T - 1) / T;
    const int NW = P.NW;
    // ---- cell order: tiles of T cells, inside a tile sub-tiles of 32 cells, inside a sub-tile ascending old ids
    std::vector<float> pos;
    detail::integrate_centres(C, Fi, owner, neigh, cellFaces, deltas, deltasUnit, pos);
    lap("cell centres");
    P.cell_new2old.resize(C);
    std::iota(P.cell_new2old.begin(), P.cell_new2old.end(), 0);
    std::vector<int> lattice;
    const int* lat = detail::detect_lattice(pos, C, lattice) ? lattice.data() : nullptr;
    lap("lattice detection");
    detail::rcb_parallel(P.cell_new2old, pos, lat, 0, C, T);
    lap("bisection into tiles");
    detail::parallel_for(P.nTiles, 256, [&](long t0, long t1) {
        for (long t = t0; t < t1; t++) {
            detail::rcb(P.cell_new2old, pos, lat, t * T, std::min<long>((t + 1) * T, C), kRound);
            for (long b = t * T; b < std::min<long>((t + 1) * T, C); b += kRound)
                std::sort(P.cell_new2old.begin() + b, P.cell_new2old.begin() + std::min<long>(b + kRound, C));
        }
    });
    lap("sub-tiles");
    P.nEarly = P.nTiles;
    if (nLocalFaces < F) {
        std::vector<char> rb(C, 0), late(P.nTiles, 0);
        for (int f = nLocalFaces; f < F; f++) rb[owner[f]] = 1;
        for (int t = 0; t < P.nTiles; t++)
            for (long c = (long)t * T; c < std::min<long>((long)(t + 1) * T, C) && !late[t]; c++) {
                const int oc = P.cell_new2old[c];
                if (rb[oc]) { late[t] = 1; break; }
                for (int j = 0; j < 6; j++) {
                    const int f = cellFaces[(size_t)oc * 6 + j];
                    if (f >= 0 && f < Fi && rb[owner[f] == oc ? neigh[f] : owner[f]]) { late[t] = 1; break; }
                }
            }
        if (C % T) late[P.nTiles - 1] = 1;           // the partial tile stays last
        std::vector<int> reordered; reordered.reserve(C);
        P.nEarly = 0;
        for (int pass = 0; pass < 2; pass++)
            for (int t = 0; t < P.nTiles; t++) {
                if (late[t] != pass) continue;
                if (!pass) P.nEarly++;
                for (long c = (long)t * T; c < std::min<long>((long)(t + 1) * T, C); c++) reordered.push_back(P.cell_new2old[c]);
            }
        P.cell_new2old.swap(reordered);
    }
    P.cell_old2new.assign(C, -1);
    for (int i = 0; i < C; i++) P.cell_old2new[P.cell_new2old[i]] = i;
    // ---- per sub-tile: entries (faces touching it) and their (round, lane) schedule. A tile's entries, schedules and halo list
    // depend on nothing outside the tile: contiguous chunks of tiles are planned on threads; the numbering of the internal faces
    // (in the order the sub-tiles first list them: a face is first listed by the tile of its lower-numbered cell) and the offsets
    // of the per-tile lists follow from prefix sums, so the plan is the same for any number of threads.
    P.face_old2new.assign(F, -1);
    P.face_new2old.assign(F, -1);
    for (int f = Fi; f < F; f++) { P.face_old2new[f] = f; P.face_new2old[f] = f; }
    const int N = C + (F - Fi);
    struct ChunkOut {
        std::vector<int> ent_face_old; std::vector<uint32_t> ent_loc; std::vector<int> halo_cell, rounds, halo_round, halo_count, firsts;
        int maxColours = 0, maxHalo = 0; long nEntries = 0;
    };
    const int nChunks = (int)std::max<long>(1, std::min<long>((long)detail::plan_threads() * 4, (P.nTiles + 63) / 64));
    std::vector<ChunkOut> outs(nChunks);
    auto chunk_lo = [&](int c) { return (int)((long)P.nTiles * c / nChunks); };
    detail::parallel_for(nChunks, 1, [&](long k0, long k1) {
        std::vector<int> faces;
        detail::SubTileSchedule sch;
        std::unordered_map<uint64_t, std::vector<detail::SubTileSchedule>> memo;
        // small open-addressing maps, cleared per tile: halo cell -> slot, internal face -> seen
        enum { kMap = 8192 };
        std::vector<int> hkey(kMap, -1), hval(kMap, 0), hused, fkey(kMap, -1), fused;
        for (long k = k0; k < k1; k++) {
            ChunkOut& o = outs[k];
            for (int t = chunk_lo((int)k); t < chunk_lo((int)k + 1); t++) {
                const int c0 = t * T, c1 = std::min(C, c0 + T);
                int nHalo = 0;
                for (int u : hused) hkey[u] = -1;
                hused.clear();
                for (int u : fused) fkey[u] = -1;
                fused.clear();
                auto slot = [&](int cell) {           // cell: NEW index (ghosts >= C)
                    if (cell >= c0 && cell < c1) return cell - c0;
                    unsigned h = ((unsigned)cell * 2654435761u) & (kMap - 1);
                    while (hkey[h] >= 0 && hkey[h] != cell) h = (h + 1) & (kMap - 1);
                    if (hkey[h] < 0) {
                        if ((int)hused.size() >= kMap / 2) throw std::runtime_error("tile halo too large");
                        hkey[h] = cell; hval[h] = T + nHalo++; hused.push_back((int)h); o.halo_cell.push_back(cell);
                    }
                    return hval[h];
                };
                auto first_seen = [&](int f) {        // true the first time this tile lists internal face f
                    unsigned h = ((unsigned)f * 2654435761u) & (kMap - 1);
                    while (fkey[h] >= 0 && fkey[h] != f) h = (h + 1) & (kMap - 1);
                    if (fkey[h] == f) return false;
                    if ((int)fused.size() >= kMap / 2) throw std::runtime_error("tile with too many faces");
                    fkey[h] = f; fused.push_back((int)h);
                    return true;
                };
                const size_t firsts0 = o.firsts.size();
                for (int w = 0; w < NW; w++) {
                    const int s0 = c0 + w * kRound, s1 = std::min(c1, s0 + kRound);
                    faces.clear(); sch.E.clear();
                    for (int c = s0; c < s1; c++) {
                        const int oc = P.cell_new2old[c];
                        for (int j = 0; j < 6; j++) {
                            const int f = cellFaces[(size_t)oc * 6 + j];
                            if (f < 0 || f >= F) throw std::runtime_error("cellFaces out of range");
                            int a = P.cell_old2new[owner[f]], b = -1;
                            if (f < Fi) {
                                b = P.cell_old2new[neigh[f]];
                                if (a == b) throw std::runtime_error("face with identical owner and neighbour");
                            }
                            const bool ain = a >= s0 && a < s1, bin = b >= s0 && b < s1;
                            // inner face of the sub-tile: list once (when reached from its lower cell)
                            if (ain && bin && c != std::min(a, b)) continue;
                            const bool atile = a >= c0 && a < c1, btile = b >= c0 && b < c1;
                            detail::SubTileSchedule::Ent e;
                            e.both = ain && bin;
                            e.a = ain ? a - s0 : b - s0; e.b = e.both ? b - s0 : -1;
                            e.pref = (atile && btile) ? 0 : 1;      // entries that read a halo slot go to the late rounds (overlap of the halo gather)
                            faces.push_back(f); sch.E.push_back(e);
                        }
                    }
                    {   // sub-tiles with the same local face graph (all interior sub-tiles of a structured block) share one schedule
                        uint64_t key = 1469598103934665603ull;
                        for (const auto& e : sch.E) { key ^= (uint64_t)(e.a | (e.b + 1) << 6 | (int)e.both << 13 | e.pref << 14); key *= 1099511628211ull; }
                        auto& bucket = memo[key];
                        const detail::SubTileSchedule* hit = nullptr;
                        for (const auto& m_ : bucket) {
                            if (m_.E.size() != sch.E.size()) continue;
                            bool same = true;
                            for (size_t i = 0; i < sch.E.size() && same; i++)
                                same = m_.E[i].a == sch.E[i].a && m_.E[i].b == sch.E[i].b && m_.E[i].both == sch.E[i].both && m_.E[i].pref == sch.E[i].pref;
                            if (same) { hit = &m_; break; }
                        }
                        if (hit) { sch.home = hit->home; sch.other = hit->other; sch.round = hit->round; sch.R = hit->R; }
                        else { sch.solve(); if (memo.size() < 8192) bucket.push_back(sch); }
                    }
                    const int Rw = faces.empty() ? 0 : sch.R;
                    o.maxColours = std::max(o.maxColours, Rw);
                    std::vector<int> at((size_t)Rw * kRound, -1), from((size_t)Rw * kRound, -1);
                    for (size_t i = 0; i < faces.size(); i++) {
                        at[(size_t)sch.round[i] * kRound + sch.home[i]] = (int)i;
                        if (sch.other[i] >= 0) from[(size_t)sch.round[i] * kRound + sch.other[i]] = sch.home[i];
                    }
                    int firstHaloRound = -1;
                    for (int r = 0; r < Rw; r++) for (int l = 0; l < kRound; l++) {
                        const int i = at[(size_t)r * kRound + l], src = from[(size_t)r * kRound + l];
                        if (i < 0) { o.ent_face_old.push_back(-1); o.ent_loc.push_back(tile_pack(0, 0, 0, 0, 0, 0, src >= 0, src >= 0 ? src : 0)); continue; }
                        const int f = faces[i];
                        const int a = P.cell_old2new[owner[f]];
                        const int b = f < Fi ? P.cell_old2new[neigh[f]] : neigh[f];
                        if (b < 0 || b >= N) throw std::runtime_error("neighbour out of range");
                        // numbered by the tile of the face's lower-numbered cell, in the order that tile first lists it
                        if (f < Fi && std::min(a, b) / T == t && first_seen(f)) o.firsts.push_back(f);
                        const bool flip = (a - s0) != l;                 // the home cell is the face's neighbour
                        if (flip && (f >= Fi || b - s0 != l)) throw std::runtime_error("tile schedule: home cell is not a cell of the face");
                        const int lo = slot(flip ? b : a), ln = slot(flip ? a : b);
                        if (lo != w * kRound + l) throw std::runtime_error("tile schedule: home slot mismatch");
                        if (ln >= 1024) throw std::runtime_error("tile halo too large");
                        if (ln >= T && firstHaloRound < 0) firstHaloRound = r;
                        o.ent_face_old.push_back(f);
                        o.ent_loc.push_back(tile_pack(ln, f < Fi ? 0 : (int)bkind[f - Fi], 1, sch.other[i] >= 0, b >= C, flip, src >= 0, src >= 0 ? src : 0));
                        o.nEntries++;
                    }
                    o.halo_round.push_back(firstHaloRound < 0 ? Rw : firstHaloRound);
                    o.rounds.push_back(Rw);
                }
                o.maxHalo = std::max(o.maxHalo, nHalo);
                o.halo_count.push_back(nHalo);
                (void)firsts0;
            }
        }
    });
    // offsets of the chunks' lists; face numbers
    std::vector<size_t> entOff(nChunks + 1, 0), haloOff(nChunks + 1, 0), faceOff(nChunks + 1, 0);
    for (int k = 0; k < nChunks; k++) {
        entOff[k + 1] = entOff[k] + outs[k].ent_face_old.size(); haloOff[k + 1] = haloOff[k] + outs[k].halo_cell.size();
        faceOff[k + 1] = faceOff[k] + outs[k].firsts.size();
        P.maxColours = std::max(P.maxColours, outs[k].maxColours); P.maxHalo = std::max(P.maxHalo, outs[k].maxHalo); P.nEntries += outs[k].nEntries;
    }
    const int nextFace = (int)faceOff[nChunks];
    if (nextFace != Fi) throw std::runtime_error("internal face not reachable from any cell (broken cellFaces)");
    detail::parallel_for(nChunks, 1, [&](long k0, long k1) {
        for (long k = k0; k < k1; k++)
            for (size_t i = 0; i < outs[k].firsts.size(); i++) {
                const int f = outs[k].firsts[i], nf = (int)(faceOff[k] + i);
                P.face_old2new[f] = nf; P.face_new2old[nf] = f;
            }
    });
    P.ent_face.resize(entOff[nChunks]); P.ent_loc.resize(entOff[nChunks]); P.halo_cell.resize(haloOff[nChunks]);
    detail::parallel_for(nChunks, 1, [&](long k0, long k1) {
        for (long k = k0; k < k1; k++) {
            const ChunkOut& o = outs[k];
            for (size_t i = 0; i < o.ent_face_old.size(); i++) {
                const int f = o.ent_face_old[i];
                P.ent_face[entOff[k] + i] = f < 0 ? -1 : P.face_old2new[f];
                P.ent_loc[entOff[k] + i] = o.ent_loc[i];
            }
            std::copy(o.halo_cell.begin(), o.halo_cell.end(), P.halo_cell.begin() + haloOff[k]);
        }
    });
    P.round_start.assign((size_t)P.nTiles * NW + 1, 0);
    P.halo_round.assign((size_t)P.nTiles * NW, 0);
    P.halo_start.assign(P.nTiles + 1, 0);
    {
        size_t sw = 0, st = 0; int rounds = 0, halo = 0;
        for (int k = 0; k < nChunks; k++) {
            for (size_t i = 0; i < outs[k].rounds.size(); i++, sw++) { P.halo_round[sw] = outs[k].halo_round[i]; rounds += outs[k].rounds[i]; P.round_start[sw + 1] = rounds; }
            for (size_t i = 0; i < outs[k].halo_count.size(); i++, st++) { halo += outs[k].halo_count[i]; P.halo_start[st + 1] = halo; }
            outs[k] = ChunkOut();             // release as we go
        }
    }
    lap("entries + schedules");
    return P;
}

}  // namespace fvm
