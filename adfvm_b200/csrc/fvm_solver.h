// Step orchestration of the residual hot path and its reverse sweep, templated on the scalar type and on an
// executor (CUDA streams/kernels in the product, plain loops in the CPU test simulator).
//
// Mirrors, per call, what the reference's generated `Function_primal` / `Function_primal_grad` do
// (call sequence verified from generated code): for each of the 3 SSPRK stages
// primitive -> halo+BC ghost fill -> [objective on stage 1] -> gradCell -> halo+BC ghost fill of gradients ->
// flux over all face classes -> [max dtc on stage 1] -> RK update; and for the adjoint the same forward
// sweep keeping every stage's state, followed by the reverse sweep stage 3 -> 1.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include "fvm_bodies.h"
#include "fvm_tile_bodies.h"
#include "fvm_tiles.h"
#include "fvm_metrics.h"
#include "fvm_viscosity.h"

namespace fvm {

static inline int pad32(long n) { return (int)(((n + 31) / 32) * 32); }

// SSPRK3 (adFVM/timestep.py:17-25)
static const double RK_ALPHA[3][3] = {{1., 0., 0.}, {3. / 4, 1. / 4, 0.}, {1. / 3, 0., 2. / 3}};
static const double RK_BETA[3] = {1., 1. / 4, 2. / 3};

struct PatchHost {
    int startFace, nFaces, cellStartFace;
    int meshType;          // 0 patch/wall-like, 1 cyclic, 2 symmetryPlane, 3 empty, 4 characteristic, 5 processor, 6 processorCyclic
    int bc[3];             // BCType for U,T,p
    int nbrPatch;          // cyclic partner index or -1
    int peer;              // processor patches: neighbour rank
    int tag;               // processor patches: ordering tag shared by both sides (adFVM/mesh.py:746-756)
};

// Exchange of processor-patch rows between ranks. send/recv are in executor memory, laid out patch-major:
// for each remote patch p (in face order) a block [ncomp][nFaces_p].
template <typename R> struct HaloComm {
    virtual ~HaloComm() {}
    virtual void exchange(const R* send, R* recv, int ncomp, const std::vector<PatchHost>& remote, void* stream) = 0;
    virtual double allreduce_sum(double v) = 0;
    // in-place sum over ranks of n scalars in executor memory, ordered on `stream` (no host synchronisation on the device)
    virtual void allreduce_sum_device(R* buf, int n, void* stream) = 0;
    virtual double allreduce_max(double v) = 0;
};

template <typename R, class Exec> class Solver {
public:
    Exec ex;
    Phys<R> ph;
    MeshDev<R> m;
    ObjDev<R> obj;
    std::vector<PatchHost> patches;          // all patches, local (sorted-name order as given) then remote
    std::vector<PatchDev<R>> patches_dev_h;  // host mirror of the device table
    HaloComm<R>* comm = nullptr;
    long launches = 0;                        // kernels launched so far (bench "gpu_launches")
    long overlap_min_cells = 1 << 20;         // adjoint_step: seeds uploaded concurrently with the forward sweep from this size on
    long graph_replays = 0;                   // steps served by replaying a captured CUDA graph (introspection)
    unsigned long long epoch = 0;             // bumped by every set_* call: part of the CUDA-graph keys (kernel arguments are baked into graphs)
    long bytes_allocated = 0;

    // device buffers
    R *W[4] = {0, 0, 0, 0}, *Q[3] = {0, 0, 0}, *G[3] = {0, 0, 0}, *S = 0, *red = 0;
    R *A[4] = {0, 0, 0, 0}, *Qb = 0, *Gb = 0, *Sb = 0;
    R *sendbuf = 0, *recvbuf = 0, *stage_aos = 0;
    int *bcells = 0; int nBcells = 0;
    TilePlan plan; int tile_cells = 128; R* tile_partial = 0; double tile_evals_per_cell = 0; int tile_colours = 0, tile_max_halo = 0;
    int tile_variant = 0;
    long tile_rounds_total = 0;               // rounds of all sub-tiles (32 lanes each; introspection)
    int nEarlyTiles = 0;                      // tiles [0,nEarlyTiles) and their cells do not depend on processor-patch data
    std::vector<int> tile_halo_hist;          // tiles per halo-size bin of 32 slots (introspection)
    enum { kHalo128r = 160, kHalo128s = 192, kHalo128 = 256, kHalo64 = 384,
           kRowSlack = 128 };   // the tile kernels bulk-copy whole T-cell rows: the last tile may read past the last row   // halo slots of the two tile-kernel instantiations (T=128: TS=384, T=64: TS=448)
    bool have_mesh = false, have_state = false, adjoint_ready = false;
    std::vector<void*> owned;                 // everything to free

    explicit Solver(const Exec& e) : ex(e) {
        std::memset(&m, 0, sizeof(m)); std::memset(&ph, 0, sizeof(ph)); std::memset(&obj, 0, sizeof(obj)); obj.kind = OBJ_NONE;
        if (const char* e = std::getenv("ADFVM_OVERLAP_MIN_CELLS")) overlap_min_cells = std::atol(e);   // tests force / forbid the overlapped path
    }
    ~Solver() { ex.sync(); for (void* p : owned) ex.free(p); ex.destroy(); }

    template <typename T> T* dalloc(size_t n) {
        T* p = (T*)ex.alloc(n * sizeof(T)); ex.zero(p, n * sizeof(T)); owned.push_back(p);
        bytes_allocated += (long)(n * sizeof(T)); return p;
    }
    template <class B> void run(int n, const B& b) { if (n > 0) { ex.run(n, b); launches++; } }

    // ---- host AoS [n][d] (reference numbering) -> device SoA [d][stride] (tile numbering: device row i = host row perm[i])
    R* upload_aos(const R* host, long n, int d, int stride, R* dst = nullptr, const int* perm = nullptr) {
        if (!dst) dst = dalloc<R>((size_t)d * stride);
        if (n == 0) return dst;
        if (d == 1 && !perm) { ex.upload(dst, host, n * sizeof(R)); return dst; }
        R* tmp = (R*)ex.alloc((size_t)n * d * sizeof(R));
        ex.upload(tmp, host, (size_t)n * d * sizeof(R));
        run((int)n, AosToSoaBody<R>{tmp, dst, d, stride, perm});
        ex.sync(); ex.free(tmp);
        return dst;
    }
    void set_physics(double gamma, double Cp, double Pr, int mu_law, double mu_value, int riemann, int briemann) {
        epoch++;
        ph.gamma = (R)gamma; ph.Cp = (R)Cp; ph.Pr = (R)Pr; ph.Cv = (R)(Cp / gamma);
        ph.gm1 = (R)(gamma - 1.); ph.iCv = (R)(gamma / Cp); ph.iCvgm1 = (R)(gamma / (Cp * (gamma - 1.)));
        ph.g_gm1 = (R)(gamma / (gamma - 1.)); ph.CpPr = (R)(Cp / Pr);
        ph.small = sizeof(R) == 8 ? (R)1e-30 : (R)1e-9;
        ph.mu_law = mu_law; ph.mu_value = (R)mu_value; ph.riemann = riemann; ph.boundary_riemann = briemann;
    }

    // ---- mesh: the 10 gradFields + 5 intFields + 8 constants of adFVM/mesh.py:27-37 as the reference passes them
    void set_mesh(const int* sizes, const R* areas, const R* volumesL, const R* volumesR, const R* weights, const R* deltas,
                  const R* normals, const R* deltasUnit, const R* linearWeights, const R* quadraticWeights, const R* volumes,
                  const int* owner, const int* neighbour, const int* cellFaces, const int* cellNeighbours, const int* cellOwner,
                  const std::vector<PatchHost>& patches_in) {
        if (have_mesh) throw std::runtime_error("mesh already set (create a new context)");
        const bool verbose = std::getenv("ADFVM_TILE_TIMING") != nullptr;
        auto t_lap = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) { if (verbose) { ex.sync(); auto now = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[set_mesh] %-26s %.2f s\n", what, std::chrono::duration<double>(now - t_lap).count()); t_lap = now; } };
        m.nCells = sizes[0]; m.nFaces = sizes[1]; m.nInternalCells = sizes[2]; m.nInternalFaces = sizes[3];
        m.nLocalCells = sizes[4]; m.nRemoteCells = sizes[5]; m.nLocalFaces = sizes[6]; m.nGhostCells = sizes[7];
        const int C = m.nInternalCells, F = m.nFaces, Fi = m.nInternalFaces, N = m.nCells;
        if (N != C + (F - Fi) || m.nGhostCells != F - Fi || m.nRemoteCells != N - m.nLocalCells ||
            m.nLocalFaces != m.nLocalCells - C + Fi)
            throw std::runtime_error("inconsistent mesh size constants");
        for (long f = 0; f < F; f++) if (owner[f] < 0 || owner[f] >= C || neighbour[f] < 0 || neighbour[f] >= N || (f < Fi && neighbour[f] >= C))
            throw std::runtime_error("owner/neighbour out of range");
        // the kernels use V[owner]/V[neighbour] for volumesL/volumesR; reject inputs where they differ
        for (long f = 0; f < F; f++) if (volumesL[f] != volumes[owner[f]]) throw std::runtime_error("volumesL != volumes[owner] (perturbed volume arrays are not supported)");
        for (long f = 0; f < Fi; f++) if (volumesR[f] != volumes[neighbour[f]]) throw std::runtime_error("volumesR != volumes[neighbour]");
        m.sC = pad32(C); m.sN = pad32(N); m.sF = pad32(F);
        lap("validation");
        // ---- tile plan: renumber internal cells and internal faces (fvm_tiles.h); ghosts and boundary faces keep their ids
        // 128-cell tiles with room for 256 halo slots; meshes with many ghost cells per cell (1-D / 2-D cases, whose
        // "empty" patches give every cell 2-4 boundary faces) fall back to 64-cell tiles, whose halo always fits
        std::vector<unsigned char> bkind(F - Fi + 1, (unsigned char)FACE_BOUNDARY);
        for (const PatchHost& h : patches_in) {
            if (h.nFaces && (h.startFace < Fi || h.startFace + h.nFaces > F)) throw std::runtime_error("patch face range out of bounds");
            const bool coupled = (h.meshType == 1 || h.meshType == 5 || h.meshType == 6);
            const unsigned char k = coupled ? FACE_COUPLED : (h.meshType == 4 ? FACE_CHARACTERISTIC : FACE_BOUNDARY);
            for (int i = 0; i < h.nFaces; i++) bkind[h.startFace - Fi + i] = k;
        }
        plan = build_tile_plan<R>(C, Fi, F, owner, neighbour, cellFaces, deltas, deltasUnit, bkind.data(), tile_cells, m.nLocalFaces);
        if (tile_cells == 128 && plan.maxHalo > kHalo128)
            plan = build_tile_plan<R>(C, Fi, F, owner, neighbour, cellFaces, deltas, deltasUnit, bkind.data(), 64, m.nLocalFaces);
        if (plan.T == 64 && plan.maxHalo > kHalo64) throw std::runtime_error("tile halo exceeds the kernel's capacity");
        // kernel variant (T, TS): 0 = (128, 288) regular 4x4x8 tiles of a hex block (halo = its 160 face neighbours: three fp64 CTAs per
        // SM forward AND reverse), 1 = (128, 320) compact 3-D tiles (halo <= 192), 2 = (128, 384), 3 = (64, 448)
        lap("tile plan");
        tile_variant = plan.T == 64 ? 3 : (plan.maxHalo <= kHalo128r ? 0 : (plan.maxHalo <= kHalo128s ? 1 : 2));
        m.T = plan.T; m.nTiles = plan.nTiles; nEarlyTiles = plan.nEarly;
        int* d_cperm = dalloc<int>(m.sC); ex.upload(d_cperm, plan.cell_new2old.data(), (size_t)C * 4); m.cell_perm = d_cperm;
        int* d_fperm = (int*)ex.alloc((size_t)(F + 1) * 4); ex.upload(d_fperm, plan.face_new2old.data(), (size_t)F * 4);
        m.area = upload_aos(areas, F, 1, m.sF, nullptr, d_fperm); m.weight = upload_aos(weights, F, 1, m.sF, nullptr, d_fperm);
        { R* idel = upload_aos(deltas, F, 1, m.sF, nullptr, d_fperm); run(F, ReciprocalBody<R>{idel}); m.idelta = idel; } m.normal = upload_aos(normals, F, 3, m.sF, nullptr, d_fperm);
        { R* d_vol = dalloc<R>((size_t)m.sC + kRowSlack); m.vol = upload_aos(volumes, C, 1, m.sC, d_vol, d_cperm); }
        for (long c = 0; c < C; c++) if (!(volumes[c] > R(0))) throw std::runtime_error("non-positive cell volume");
        {   // per-pass metric chunks of the flux kernels, gathered on the device from temporary face-indexed arrays
            R* t_dunit = (R*)ex.alloc((size_t)3 * m.sF * sizeof(R)); upload_aos(deltasUnit, F, 3, m.sF, t_dunit, d_fperm);
            R* t_linw = (R*)ex.alloc((size_t)2 * m.sF * sizeof(R)); upload_aos(linearWeights, F, 2, m.sF, t_linw, d_fperm);
            R* t_quadw = (R*)ex.alloc((size_t)6 * m.sF * sizeof(R)); upload_aos(quadraticWeights, F, 6, m.sF, t_quadw, d_fperm);
            const size_t nslots = plan.ent_face.size();
            int* t_face = (int*)ex.alloc((nslots + 1) * 4); ex.upload(t_face, plan.ent_face.data(), nslots * 4);
            unsigned* t_word = (unsigned*)ex.alloc((nslots + 1) * 4); ex.upload(t_word, plan.ent_loc.data(), nslots * 4);
            R* d_chunks = dalloc<R>((nslots / kRound) * (size_t)Chunk<R, kRound>::kScalars + 16);
            run((int)nslots, FillChunksBody<R, kRound>{t_face, t_word, m.sF, m.area, m.normal, m.idelta, t_dunit, t_linw, t_quadw, d_chunks});
            m.chunks = d_chunks;
            ex.sync(); ex.free(t_face); ex.free(t_word);
            if (mesh_param) { f_dunit = t_dunit; f_linw = t_linw; f_quadw = t_quadw; owned.push_back(t_dunit); owned.push_back(t_linw); owned.push_back(t_quadw); }
            else { ex.free(t_dunit); ex.free(t_linw); ex.free(t_quadw); }
        }
        ex.sync(); ex.free(d_fperm);
        lap("metric upload + chunks");
        auto newcell = [&](int c) { return c < C ? plan.cell_old2new[c] : c; };
        {
            std::vector<int> ow(m.sF, 0), nb(m.sF, 0);
            detail::parallel_for(F, 1 << 16, [&](long f_lo, long f_hi) {
                for (long f = f_lo; f < f_hi; f++) {
                    const int of = plan.face_new2old[f];
                    if (owner[of] < 0 || owner[of] >= C || neighbour[of] < 0 || neighbour[of] >= N) throw std::runtime_error("owner/neighbour out of range");
                    ow[f] = newcell(owner[of]); nb[f] = newcell(neighbour[of]);
                }
            });
            int* d_owner = dalloc<int>(m.sF); ex.upload(d_owner, ow.data(), (size_t)F * 4); m.owner = d_owner;
            int* d_neigh = dalloc<int>(m.sF); ex.upload(d_neigh, nb.data(), (size_t)F * 4); m.neigh = d_neigh;
            ex.sync();
        }
        // connectivity: transpose + renumber on the host (ints, one-off)
        std::vector<int> cf((size_t)6 * m.sC, 0), cn((size_t)6 * m.sC, 0);
        std::vector<unsigned char> co(m.sC, 0);
        std::vector<int> bc_list;
        {
            std::vector<unsigned char> isb(C, 0);
            detail::parallel_for(C, 1 << 16, [&](long c_lo, long c_hi) {
                for (long c = c_lo; c < c_hi; c++) {
                    const long oc = plan.cell_new2old[c];
                    bool b = false; unsigned bits = 0;
                    for (int j = 0; j < 6; j++) {
                        int f = cellFaces[oc * 6 + j], nb = cellNeighbours[oc * 6 + j];
                        if (f < 0 || f >= F || nb < 0 || nb >= N) throw std::runtime_error("cellFaces/cellNeighbours out of range");
                        cf[(size_t)j * m.sC + c] = plan.face_old2new[f]; cn[(size_t)j * m.sC + c] = newcell(nb);
                        if (cellOwner[oc * 6 + j]) bits |= 1u << j;
                        if (nb >= C) b = true;
                    }
                    co[c] = (unsigned char)bits;
                    isb[c] = b;
                }
            });
            for (long c = 0; c < C; c++) if (isb[c]) bc_list.push_back((int)c);
        }
        lap("connectivity transpose");
        int* d_cf = dalloc<int>(cf.size()); ex.upload(d_cf, cf.data(), cf.size() * 4); m.cellFaces = d_cf;
        int* d_cn = dalloc<int>(cn.size()); ex.upload(d_cn, cn.data(), cn.size() * 4); m.cellNbr = d_cn;
        unsigned char* d_co = dalloc<unsigned char>(co.size()); ex.upload(d_co, co.data(), co.size()); m.cellOwner = d_co;
        { R* d_cfm = dalloc<R>((size_t)24 * m.sC); run(C, CellFaceMetricBody<R>{m, d_cfm}); m.cfm = d_cfm; }
        nBcells = (int)bc_list.size();
        bcells = dalloc<int>(nBcells + 1); ex.upload(bcells, bc_list.data(), (size_t)nBcells * 4);
        {
            int* d_ps = dalloc<int>(plan.round_start.size()); ex.upload(d_ps, plan.round_start.data(), plan.round_start.size() * 4); m.round_start = d_ps;
            int* d_hp = dalloc<int>(plan.halo_round.size() + 1); ex.upload(d_hp, plan.halo_round.data(), plan.halo_round.size() * 4); m.halo_round = d_hp;
            int* d_hs = dalloc<int>(plan.halo_start.size()); ex.upload(d_hs, plan.halo_start.data(), plan.halo_start.size() * 4); m.halo_start = d_hs;
            int* d_hc = dalloc<int>(plan.halo_cell.size() + 1); ex.upload(d_hc, plan.halo_cell.data(), plan.halo_cell.size() * 4); m.halo_cell = d_hc;
            tile_partial = dalloc<R>((size_t)plan.nTiles * plan.NW + 1);
            unsigned short* d_ns = dalloc<unsigned short>((size_t)6 * m.sC); run(C, NbrSlotBody<R>{m, d_ns}); m.nbrSlot = d_ns;
            ex.sync();
        }
        lap("connectivity upload + slots");
        if (mesh_param) { face_new2old_h = plan.face_new2old; cell_new2old_h = plan.cell_new2old; }
        tile_evals_per_cell = plan.evals_per_cell(); tile_colours = plan.maxColours; tile_max_halo = plan.maxHalo;
        tile_rounds_total = (long)(plan.ent_face.size() / kRound);
        tile_halo_hist.assign(32, 0);
        for (int t = 0; t < plan.nTiles; t++) tile_halo_hist[std::min(31, (plan.halo_start[t + 1] - plan.halo_start[t]) / 32)]++;
        // the plan's host vectors are only needed for the I/O permutation, which lives on the device: release them
        plan = TilePlan(); plan.T = m.T; plan.nTiles = m.nTiles;
        // patches
        patches = patches_in;
        if ((int)patches.size() > MAX_PATCHES) throw std::runtime_error("too many patches");
        std::vector<unsigned char> bp(m.nGhostCells + 1, 255);
        patches_dev_h.assign(patches.size(), PatchDev<R>());
        for (size_t p = 0; p < patches.size(); p++) {
            const PatchHost& h = patches[p]; PatchDev<R>& d = patches_dev_h[p];
            std::memset(&d, 0, sizeof(d));
            d.startFace = h.startFace; d.nFaces = h.nFaces; d.cellStartFace = h.cellStartFace;
            if (h.nFaces && (h.startFace < Fi || h.startFace + h.nFaces > F || h.cellStartFace != h.startFace - Fi + C))
                throw std::runtime_error("patch face range out of bounds");
            const bool coupled = (h.meshType == 1 || h.meshType == 5 || h.meshType == 6);
            d.kind = coupled ? FACE_COUPLED : (h.meshType == 4 ? FACE_CHARACTERISTIC : FACE_BOUNDARY);
            for (int k = 0; k < 3; k++) d.bc[k] = h.bc[k];
            d.gbc = (h.meshType == 1) ? BC_CYCLIC : ((h.meshType == 5 || h.meshType == 6) ? BC_PROCESSOR : BC_ZEROGRADIENT);
            if (h.meshType == 1) {
                if (h.nbrPatch < 0 || h.nbrPatch >= (int)patches.size() || patches[h.nbrPatch].nFaces != h.nFaces)
                    throw std::runtime_error("cyclic patch without a matching neighbourPatch");
                d.nbrStartFace = patches[h.nbrPatch].startFace; d.nbrCellStartFace = patches[h.nbrPatch].cellStartFace;
            }
            for (int i = 0; i < h.nFaces; i++) bp[h.startFace - Fi + i] = (unsigned char)p;
        }
        for (int b = 0; b < m.nGhostCells; b++) if (bp[b] == 255) throw std::runtime_error("boundary face not covered by any patch");
        unsigned char* d_bp = dalloc<unsigned char>(bp.size()); ex.upload(d_bp, bp.data(), bp.size()); m.bpatch = d_bp;
        m.nPatches = (int)patches.size();
        patches_dev = dalloc<PatchDev<R>>(patches.size() + 1);
        m.patches = patches_dev;
        push_patches();
        // state + work buffers
        for (int k = 0; k < 4; k++) W[k] = dalloc<R>((size_t)5 * m.sC + kRowSlack);
        for (int k = 0; k < 2; k++) Q[k] = dalloc<R>((size_t)5 * m.sN + kRowSlack);
        G[0] = dalloc<R>((size_t)15 * m.sN + kRowSlack);
        S = dalloc<R>((size_t)5 * m.sC + kRowSlack);
        red = dalloc<R>(8);
        stage_aos = dalloc<R>((size_t)5 * m.sC);
        if (m.nRemoteCells > 0) { sendbuf = dalloc<R>((size_t)15 * m.nRemoteCells); recvbuf = dalloc<R>((size_t)15 * m.nRemoteCells); }
        ex.sync();
        lap("patches + buffers");
        have_mesh = true;
    }
    PatchDev<R>* patches_dev = nullptr;
    void push_patches() { ex.upload(patches_dev, patches_dev_h.data(), patches_dev_h.size() * sizeof(PatchDev<R>)); ex.sync(); }

    // BC value arrays (static inputs created by BoundaryCondition.createInput, adFVM/BCs.py:45-54); host AoS [nFaces][d]
    enum BCKey { KEY_VALUE_U = 0, KEY_VALUE_T, KEY_VALUE_P, KEY_U0, KEY_T0, KEY_P0, KEY_TT, KEY_PT, KEY_DIR };
    void set_bc_value(int patch, int key, const R* host) {
        epoch++;
        if (patch < 0 || patch >= (int)patches.size()) throw std::runtime_error("bad patch index");
        const int n = patches[patch].nFaces;
        const int d = (key == KEY_VALUE_U || key == KEY_U0 || key == KEY_DIR) ? 3 : 1;
        PatchDev<R>& pd = patches_dev_h[patch];
        const R** slot;
        switch (key) {
        case KEY_VALUE_U: slot = &pd.valU; break; case KEY_VALUE_T: slot = &pd.valT; break; case KEY_VALUE_P: slot = &pd.valp; break;
        case KEY_U0: slot = &pd.U0; break; case KEY_T0: slot = &pd.T0; break; case KEY_P0: slot = &pd.p0; break;
        case KEY_TT: slot = &pd.Tt; break; case KEY_PT: slot = &pd.pt; break; case KEY_DIR: slot = &pd.dir; break;
        default: throw std::runtime_error("bad BC key");
        }
        R* dst = const_cast<R*>(*slot);
        dst = upload_aos(host, n, d, n > 0 ? n : 1, dst);
        *slot = dst;
        push_patches();
    }
    // parameter block of the adjoint: the source terms (default) or one BC input array (reference apps/adjoint.py:101-120)
    int param_patch = -1, param_key = -1, param_dim = 0; R* Pb = nullptr;
    // parameters = 'mesh': must be chosen before set_mesh (the face-indexed flux metrics and the numbering maps are kept)
    bool mesh_param = false; R *Mb = nullptr, *Vb = nullptr, *f_dunit = nullptr, *f_linw = nullptr, *f_quadw = nullptr;
    std::vector<int> face_new2old_h, cell_new2old_h;
    void set_parameter_mesh() { if (have_mesh) throw std::runtime_error("select parameters='mesh' before the mesh is set"); mesh_param = true; epoch++; }
    void set_parameter_source() { epoch++; param_patch = -1; param_key = -1; }
    void set_parameter_bc(int patch, int key) {
        epoch++;
        if (patch < 0 || patch >= (int)patches.size()) throw std::runtime_error("bad patch index");
        const PatchDev<R>& d = patches_dev_h[patch];
        const bool ok = (key == KEY_VALUE_U && d.bc[0] == BC_FIXEDVALUE) || (key == KEY_VALUE_T && d.bc[1] == BC_FIXEDVALUE) ||
                        (key == KEY_VALUE_P && d.bc[2] == BC_FIXEDVALUE) ||
                        ((key == KEY_U0 || key == KEY_T0 || key == KEY_P0) && d.bc[2] == BC_CBC_UPT) ||
                        ((key == KEY_TT || key == KEY_PT) && d.bc[2] == BC_CBC_TOTAL_PT);
        if (!ok) throw std::runtime_error("this boundary condition has no such input to differentiate (or it is not supported: direction)");
        param_patch = patch; param_key = key; param_dim = (key == KEY_VALUE_U || key == KEY_U0) ? 3 : 1;
        Pb = dalloc<R>((size_t)param_dim * std::max(1, patches[patch].nFaces));
    }
    // host [nFaces][d] <- device [d][nFaces]
    void get_param_grad(R* out, bool zero_after) {
        if (param_patch < 0) throw std::runtime_error("no BC parameter selected");
        const int n = patches[param_patch].nFaces, d = param_dim;
        std::vector<R> h((size_t)n * d);
        ex.download(h.data(), Pb, h.size() * sizeof(R)); ex.sync();
        for (int i = 0; i < n; i++) for (int k = 0; k < d; k++) out[(size_t)i * d + k] = h[(size_t)k * n + i];
        if (zero_after) ex.zero(Pb, h.size() * sizeof(R));
    }
    // the ten gradient arrays in the reference's layout and numbering (adFVM/mesh.py:27-31): areas [F], volumesL [F],
    // volumesR [Fi], weights [F], deltas [F], normals [F][3], deltasUnit [F][3], linearWeights [F][2], quadraticWeights [F][2][3],
    // volumes [C]
    void get_mesh_grad(R* const out[10], bool zero_after) {
        if (!mesh_param || !Mb) throw std::runtime_error("parameters='mesh' was not selected (or no adjoint step has run)");
        const int F = m.nFaces, Fi = m.nInternalFaces, C = m.nInternalCells;
        std::vector<R> h((size_t)19 * m.sF), v((size_t)m.sC);
        ex.download(h.data(), Mb, h.size() * sizeof(R)); ex.download(v.data(), Vb, v.size() * sizeof(R)); ex.sync();
        auto row = [&](int r, long f) { return h[(size_t)r * m.sF + f]; };
        for (long f = 0; f < F; f++) {
            const long of = face_new2old_h[f];
            out[0][of] = row(0, f); out[1][of] = row(1, f); if (of < Fi) out[2][of] = row(2, f);
            out[3][of] = row(3, f); out[4][of] = row(4, f);
            for (int k = 0; k < 3; k++) { out[5][of * 3 + k] = row(5 + k, f); out[6][of * 3 + k] = row(8 + k, f); }
            out[7][of * 2] = row(11, f); out[7][of * 2 + 1] = row(12, f);
            for (int k = 0; k < 6; k++) out[8][of * 6 + k] = row(13 + k, f);
        }
        for (long c = 0; c < C; c++) out[9][cell_new2old_h[c]] = v[c];
        if (zero_after) { ex.zero(Mb, h.size() * sizeof(R)); ex.zero(Vb, v.size() * sizeof(R)); }
    }
    void check_bcs() const {
        for (size_t p = 0; p < patches.size(); p++) {
            const PatchDev<R>& d = patches_dev_h[p];
            if (d.nFaces == 0) continue;
            if ((d.bc[0] == BC_FIXEDVALUE && !d.valU) || (d.bc[1] == BC_FIXEDVALUE && !d.valT) || (d.bc[2] == BC_FIXEDVALUE && !d.valp) ||
                (d.bc[2] == BC_CBC_UPT && !(d.U0 && d.T0 && d.p0)) || (d.bc[2] == BC_CBC_TOTAL_PT && !(d.Tt && d.pt)))
                throw std::runtime_error("boundary condition input array missing for patch " + std::to_string(p));
        }
    }
    void set_objective(int kind, int patch, int dir) {
        epoch++;
        if (kind != OBJ_NONE && kind != OBJ_CELL_TV && kind != OBJ_CELL_T && (patch < 0 || patch >= (int)patches.size())) throw std::runtime_error("objective patch out of range");
        obj.kind = kind; obj.patch = patch; obj.dir = dir;
    }
    // cut-plane objective (reference adFVM/objectives/vane.py): cells in the reference's numbering, areas, constants
    void set_objective_plane(int n, const int* cells, const R* areas, double ptin, const double* normal, double scale) {
        epoch++;
        if (!have_mesh) throw std::runtime_error("set the mesh before the objective");
        const int C = m.nInternalCells;
        for (int i = 0; i < n; i++) if (cells[i] < 0 || cells[i] >= C) throw std::runtime_error("objective plane cell out of range");
        int* inv = (int*)ex.alloc((size_t)(C + 1) * 4);
        int* d_in = (int*)ex.alloc((size_t)(n + 1) * 4);
        int* d_cells = dalloc<int>(n + 1);
        R* d_areas = dalloc<R>(n + 1);
        ex.upload(d_in, cells, (size_t)n * 4); ex.upload(d_areas, areas, (size_t)n * sizeof(R));
        run(C, InvertPermBody<R>{m.cell_perm, inv});
        run(n, GatherIntBody<R>{inv, d_in, d_cells});
        ex.sync(); ex.free(inv); ex.free(d_in);
        obj.kind = OBJ_PLANE_PTLOSS; obj.patch = 0; obj.dir = 0; obj.cells = d_cells; obj.areas = d_areas; obj.ncells = n;
        obj.ptin = (R)ptin; obj.scale = (R)scale;
        for (int k = 0; k < 3; k++) obj.nrm[k] = (R)normal[k];
    }
    // objective evaluated by the host layer (any case-file objective, traced by the reference's own front-end)
    typedef double (*ObjectiveFn)(void* user, const void* Q, long long stride, int want_seed, double obja, void* Qseed);
    ObjectiveFn obj_fn = nullptr; void* obj_user = nullptr; R* Qseed = nullptr; R obj_host = R(0);
    void set_objective_callback(ObjectiveFn fn, void* user) {
        epoch++;
        if (!fn) throw std::runtime_error("null objective callback");
        obj_fn = fn; obj_user = user; obj.kind = OBJ_CALLBACK; obj.patch = 0; obj.dir = 0;
    }
    void get_cell_perm(int* out) { ex.download(out, m.cell_perm, (size_t)m.nInternalCells * 4); ex.sync(); }
    void objective_forward(const R* Qs) {            // -> red[1] (red[2] = plane mass flux)
        const int C = m.nInternalCells;
        if (obj.kind == OBJ_NONE) { ex.zero(red + 1, sizeof(R)); return; }
        if (obj.kind == OBJ_CALLBACK) {
            obj_host = (R)obj_fn(obj_user, Qs, (long long)m.sN, 0, 0., nullptr);
            ex.upload(red + 1, &obj_host, sizeof(R));
            return;
        }
        if (obj.kind == OBJ_PLANE_PTLOSS) {
            // the mass flux of the whole plane normalises every rank's share (mpi_allreduce of w, objectives/vane.py:97-99);
            // the shares themselves are summed over the ranks with the objective (get_dtc_obj)
            ex.reduce_sum(obj.ncells, PlaneMassBody<R>{ph, m, obj, Qs}, red + 2);
            if (comm) comm->allreduce_sum_device(red + 2, 1, ex.stream_handle());
            ex.reduce_sum(obj.ncells, PlaneLossBody<R>{ph, m, obj, Qs, red + 2}, red + 1); launches += 4;
            return;
        }
        const int n = (obj.kind == OBJ_CELL_TV || obj.kind == OBJ_CELL_T) ? C : patches[obj.patch].nFaces;
        ex.reduce_sum(n, ObjectiveBody<R>{ph, m, obj, Qs}, red + 1); launches += 2;
    }
    void set_source(const R* s_rho, const R* s_rhoU, const R* s_rhoE) {
        const int C = m.nInternalCells;
        upload_aos(s_rho, C, 1, m.sC, S, m.cell_perm); upload_aos(s_rhoU, C, 3, m.sC, S + m.sC, m.cell_perm);
        upload_aos(s_rhoE, C, 1, m.sC, S + 4 * (size_t)m.sC, m.cell_perm);
    }
    void set_state(const R* rho, const R* rhoU, const R* rhoE) { put5(W[0], rho, rhoU, rhoE); have_state = true; }
    // primal_grad from host arrays must not disturb the resident state of `primal` (the reference keeps that one under the
    // reuse ids primal_0..2 of Function_primal, adpy/adpy/variable.py:382-388): the adjoint's start state goes to a spare
    // buffer that takes the place of W[0] for the duration of the call
    R* Wspare = nullptr; bool adj_had_state = false;
    void adjoint_state_begin(const R* rho, const R* rhoU, const R* rhoE) {
        adj_had_state = have_state;
        if (have_state) { if (!Wspare) Wspare = dalloc<R>((size_t)5 * m.sC + kRowSlack); std::swap(W[0], Wspare); }
        put5(W[0], rho, rhoU, rhoE);
        have_state = true;
    }
    void adjoint_state_end() { if (adj_had_state) std::swap(W[0], Wspare); have_state = adj_had_state; }
    // ---- opt-in cache of states handed out to the host (adfvm_state_cache_*): a state the host layer can PROVE unchanged (the very
    // array objects it returned, still read-only) is taken from its device copy instead of travelling up again. The reference
    // re-uploads it at every `primal_grad` call (apps/adjoint.py:268-280: no reuse ids on that function).
    struct CacheSlot { R* buf; long long key; long stamp; };
    std::vector<CacheSlot> scache; int scache_cap = 0; long scache_clock = 0, scache_hits = 0; R* scache_pending = nullptr;
    void state_cache_reserve(int n) { if (n < 0) throw std::runtime_error("bad cache size"); scache_cap = n; }
    void state_cache_put(long long key) {             // the resident state W[0] under `key`
        if (scache_cap == 0 || !have_state) return;
        CacheSlot* s = nullptr;
        for (auto& c : scache) if (c.key == key) s = &c;
        if (!s && (int)scache.size() < scache_cap) { scache.push_back({dalloc<R>((size_t)5 * m.sC + kRowSlack), key, 0}); s = &scache.back(); }
        if (!s) { s = &scache[0]; for (auto& c : scache) if (c.stamp < s->stamp) s = &c; }          // least recently used
        s->key = key; s->stamp = ++scache_clock;
        ex.copy(s->buf, W[0], (size_t)5 * m.sC * sizeof(R));
    }
    bool state_cache_select(long long key) {          // the next call whose state arrays are NULL takes this copy
        scache_pending = nullptr;
        for (auto& c : scache) if (c.key == key) { c.stamp = ++scache_clock; scache_pending = c.buf; }
        return scache_pending != nullptr;
    }
    R* state_cache_take() {
        if (!scache_pending) throw std::runtime_error("state arrays required (no cached state selected)");
        R* b = scache_pending; scache_pending = nullptr; scache_hits++;
        return b;
    }
    void set_state_cached() { ex.copy(W[0], state_cache_take(), (size_t)5 * m.sC * sizeof(R)); have_state = true; }
    void adjoint_state_begin_cached() {
        R* src = state_cache_take();
        adj_had_state = have_state;
        if (have_state) { if (!Wspare) Wspare = dalloc<R>((size_t)5 * m.sC + kRowSlack); std::swap(W[0], Wspare); }
        ex.copy(W[0], src, (size_t)5 * m.sC * sizeof(R));
        have_state = true;
    }
    void get_state(R* rho, R* rhoU, R* rhoE) { get5(W[0], rho, rhoU, rhoE); }
    // host (rho[C][1], rhoU[C][3], rhoE[C][1]) in reference cell order <-> device [5][sC] in tile order
    void put5(R* dst, const R* a, const R* b, const R* c) {
        const int C = m.nInternalCells;
        ex.upload(stage_aos, a, (size_t)C * sizeof(R));
        ex.upload(stage_aos + C, b, (size_t)3 * C * sizeof(R));
        ex.upload(stage_aos + 4 * (size_t)C, c, (size_t)C * sizeof(R));
        run(C, AosToSoaBody<R>{stage_aos, dst, 1, m.sC, m.cell_perm});
        run(C, AosToSoaBody<R>{stage_aos + C, dst + m.sC, 3, m.sC, m.cell_perm});
        run(C, AosToSoaBody<R>{stage_aos + 4 * (size_t)C, dst + 4 * (size_t)m.sC, 1, m.sC, m.cell_perm});
    }
    void get5(const R* src, R* a, R* b, R* c) {
        const int C = m.nInternalCells;
        run(C, SoaToAosBody<R>{src, stage_aos, 1, m.sC, m.cell_perm});
        run(C, SoaToAosBody<R>{src + m.sC, stage_aos + C, 3, m.sC, m.cell_perm});
        run(C, SoaToAosBody<R>{src + 4 * (size_t)m.sC, stage_aos + 4 * (size_t)C, 1, m.sC, m.cell_perm});
        ex.download(a, stage_aos, (size_t)C * sizeof(R));
        ex.download(b, stage_aos + C, (size_t)3 * C * sizeof(R));
        ex.download(c, stage_aos + 4 * (size_t)C, (size_t)C * sizeof(R));
        ex.sync();
    }

    // the reference's `init` function (Function('init', ...), adFVM/density.py:64-80): primitive -> ghost fill (BCs + halo) ->
    // conservative on ALL rows; host arrays [nCells][d] out (reference numbering). Does not touch the resident state.
    void init_fields(const R* rho, const R* rhoU, const R* rhoE, R* orho, R* orhoU, R* orhoE) {
        if (!have_mesh) throw std::runtime_error("mesh not set");
        check_bcs();
        const int C = m.nInternalCells, N = m.nCells, nLB = m.nLocalFaces - m.nInternalFaces;
        R* Wt = (R*)ex.alloc(((size_t)5 * m.sC + kRowSlack) * sizeof(R));
        R* out = (R*)ex.alloc((size_t)5 * m.sN * sizeof(R));
        R* aos = (R*)ex.alloc((size_t)5 * N * sizeof(R));
        put5(Wt, rho, rhoU, rhoE);
        run(C, PrimitiveBody<R>{ph, m.sC, m.sN, Wt, Q[0]});
        run(nLB, GhostPrimBody<R>{ph, m, Q[0]});
        halo_begin(Q[0], 5); halo_end();
        run(N, ConservativeAllBody<R>{ph, m.sN, Q[0], out});
        const int dims[3] = {1, 3, 1}; R* host[3] = {orho, orhoU, orhoE};
        size_t off = 0; int row = 0;
        for (int f = 0; f < 3; f++) {
            run(C, SoaToAosBody<R>{out + (size_t)row * m.sN, aos + off, dims[f], m.sN, m.cell_perm});
            run(N - C, SoaToAosBody<R>{out + (size_t)row * m.sN + C, aos + off + (size_t)C * dims[f], dims[f], m.sN, nullptr});
            ex.download(host[f], aos + off, (size_t)N * dims[f] * sizeof(R));
            off += (size_t)N * dims[f]; row += dims[f];
        }
        ex.sync();
        ex.free(Wt); ex.free(out); ex.free(aos);
    }
    // ---- processor-patch halo (replaces adFVM/cpp/parallel.cpp Function_mpi_init/mpi/mpi_end)
    std::vector<PatchHost> remote_patches() const {
        std::vector<PatchHost> r;
        for (const PatchHost& p : patches) if (p.meshType == 5 || p.meshType == 6) r.push_back(p);
        return r;
    }
    // The exchange runs on the executor's side stream between halo_begin and halo_end; kernels issued on the main
    // stream in between overlap with it (they must not touch remote ghost rows, sendbuf or recvbuf).
    void halo_begin(R* X, int ncomp) {
        if (m.nRemoteCells == 0) return;
        if (!comm) throw std::runtime_error("mesh has processor patches but no communicator was attached");
        run(m.nRemoteCells, HaloPackBody<R>{m, X, ncomp, sendbuf});
        ex.side_begin();
        comm->exchange(sendbuf, recvbuf, ncomp, remote_patches(), ex.stream_handle());
        run(m.nRemoteCells, HaloUnpackBody<R>{m, X, ncomp, recvbuf});
        ex.side_end();
    }
    void halo_end() { if (m.nRemoteCells > 0) ex.join(); }
    void halo_reverse_begin(const R* Xb, int ncomp) {
        if (m.nRemoteCells == 0) return;
        if (!comm) throw std::runtime_error("mesh has processor patches but no communicator was attached");
        run(m.nRemoteCells, HaloPackGhostBody<R>{m, Xb, ncomp, sendbuf});
        ex.side_begin();
        comm->exchange(sendbuf, recvbuf, ncomp, remote_patches(), ex.stream_handle());
        ex.side_end();
    }
    const R* halo_reverse_end() { if (m.nRemoteCells == 0) return nullptr; ex.join(); return recvbuf; }
    // cells / tiles split for the overlap: [0, early) while the halo is in flight, the rest after it
    int early_tiles() const { return m.nRemoteCells > 0 ? nEarlyTiles : m.nTiles; }
    int early_cells() const { const long e = (long)early_tiles() * m.T; return (int)(e < m.nInternalCells ? e : m.nInternalCells); }
    template <class B> void run_range(int first, int n, const B& b) { if (n > 0) { ex.run_range(first, n, b); launches++; } }
    template <class B> void run_tiles_range(int first, int n, const B& b) { if (n > 0) { ex.run_tiles_range(first, n, b); launches++; } }

    // ---- one residual stage + RK update
    // flux=false: stop after the gradients (the adjoint's forward sweep needs Q,G of the last stage, not its output state)
    void stage(int s, R dt, R* Qs, R* Gs, R* Qnext, bool want_dtc_obj, bool flux = true) {
        const int C = m.nInternalCells, nLB = m.nLocalFaces - m.nInternalFaces;
        if (s == 0) run(C, PrimitiveBody<R>{ph, m.sC, m.sN, W[0], Qs});
        const int Ce = early_cells(), Te = early_tiles();
        run(nLB, GhostPrimBody<R>{ph, m, Qs});
        halo_begin(Qs, 5);
        if (want_dtc_obj) objective_forward(Qs);
        run_range(0, Ce, GradCellBody<R>{m, Qs, Gs});                  // overlaps the exchange of U,T,p
        halo_end();
        run_range(Ce, C - Ce, GradCellBody<R>{m, Qs, Gs});
        run(nLB, GhostGradBody<R>{m, Gs});
        halo_begin(Gs, 15);
        if (!flux) { halo_end(); return; }
        for (int part = 0; part < 2; part++) {                          // early tiles overlap the exchange of the gradients
            const int t0 = part ? Te : 0, nt = part ? m.nTiles - Te : Te;
            if (part) halo_end();
            if (tile_variant == 0) run_flux_tile<128, 128 + kHalo128r>(s, dt, Qs, Gs, Qnext, want_dtc_obj, t0, nt);
            else if (tile_variant == 1) run_flux_tile<128, 128 + kHalo128s>(s, dt, Qs, Gs, Qnext, want_dtc_obj, t0, nt);
            else if (tile_variant == 2) run_flux_tile<128, 128 + kHalo128>(s, dt, Qs, Gs, Qnext, want_dtc_obj, t0, nt);
            else run_flux_tile<64, 64 + kHalo64>(s, dt, Qs, Gs, Qnext, want_dtc_obj, t0, nt);
        }
        if (want_dtc_obj) { ex.reduce_max_buffer(tile_partial, m.nTiles * (m.T / kRound), red); launches += 2; }
    }

    template <int T, int TS> void run_flux_tile(int s, R dt, R* Qs, R* Gs, R* Qnext, bool want_dtc_obj, int t0, int nt) {
        FluxTileBody<R, T, TS> fb;
        fb.ph = ph; fb.m = m; fb.Q = Qs; fb.G = Gs;
        fb.W0 = W[0]; fb.W1 = RK_ALPHA[s][1] != 0. ? W[1] : nullptr; fb.W2 = RK_ALPHA[s][2] != 0. ? W[2] : nullptr;
        fb.a0 = (R)RK_ALPHA[s][0]; fb.a1 = (R)RK_ALPHA[s][1]; fb.a2 = (R)RK_ALPHA[s][2];
        fb.beta = (R)RK_BETA[s]; fb.dt = dt; fb.S = S; fb.Wn = W[s + 1]; fb.Qn = Qnext;
        fb.dtc_partial = want_dtc_obj ? tile_partial : nullptr;
        run_tiles_range(t0, nt, fb);
    }

    // runs `body` (a whole step: kernel launches on the executor's stream only, no host synchronisation) as a CUDA graph
    // keyed by everything the launches depend on; eager on the first occurrence of a key and wherever graphs do not apply
    // (multi-rank steps use a second stream and NCCL, per-kernel timing records events)
    static unsigned long long bits(double v) { unsigned long long u; std::memcpy(&u, &v, 8); return u; }
    template <class F> void with_graph(const std::vector<unsigned long long>& key, F&& body) {
        if (!ex.graph_usable() || comm || m.nRemoteCells > 0 || obj.kind == OBJ_CALLBACK) { body(); return; }   // (host callbacks cannot be captured)
        long n = 0;
        if (ex.graph_launch(key, n)) { launches += n; graph_replays++; return; }
        if (ex.graph_first_time(key) || !ex.graph_begin()) { body(); return; }
        const long l0 = launches;
        body();
        if (!ex.graph_end_launch(key, launches - l0)) { launches = l0; body(); }
    }

    // primal step; state W[0] -> W[0]. keep=true keeps every stage (Q[s], G[s], W[s]) for the reverse sweep.
    void primal_step(R dt, bool keep = false) {
        if (!have_mesh || !have_state) throw std::runtime_error("mesh/state not set");
        check_bcs();
        if (keep) ensure_adjoint_buffers();
        auto body = [&]() {
            for (int s = 0; s < 3; s++) {
                R* Qs = keep ? Q[s] : Q[s % 2];
                R* Gs = keep ? G[s] : G[0];
                R* Qn = (s < 2) ? (keep ? Q[s + 1] : Q[(s + 1) % 2]) : nullptr;
                stage(s, dt, Qs, Gs, Qn, s == 1 && !keep, !(keep && s == 2));
            }
        };
        if (keep) body();                          // part of the adjoint step's graph
        else with_graph({1ull, epoch, bits((double)dt), (unsigned long long)W[0], (unsigned long long)W[3], (unsigned long long)obj.kind,
                         (unsigned long long)obj.patch, (unsigned long long)obj.dir, (unsigned long long)patches_dev, (unsigned long long)obj.cells}, body);
        if (!keep) { R* t = W[0]; W[0] = W[3]; W[3] = t; steps_done++; }
    }
    // dtc (max over ranks not applied here: the reference returns the rank-local max, adFVM/density.py:405-413; the global one
    // is adfvm_get_dtc_global) and objective (allreduce-summed like mpi_allreduce, adFVM/cpp/parallel.cpp:214-232)
    double obj_cached = 0.; long obj_cached_step = -1, steps_done = 0;
    void get_dtc_obj(double* dtc, double* objective) {
        R h[2]; ex.download(h, red, 2 * sizeof(R)); ex.sync();
        *dtc = (double)h[0];
        if (obj_cached_step != steps_done) {           // one collective per step, however often the pair is read
            obj_cached = comm ? comm->allreduce_sum((double)h[1]) : (double)h[1];
            obj_cached_step = steps_done;
        }
        *objective = obj_cached;
    }

    void ensure_adjoint_buffers() {
        if (adjoint_ready) return;
        Q[2] = dalloc<R>((size_t)5 * m.sN + kRowSlack);
        G[1] = dalloc<R>((size_t)15 * m.sN + kRowSlack); G[2] = dalloc<R>((size_t)15 * m.sN + kRowSlack);
        for (int k = 0; k < 4; k++) A[k] = dalloc<R>((size_t)5 * m.sC + kRowSlack);
        Qb = dalloc<R>((size_t)5 * m.sN); Gb = dalloc<R>((size_t)15 * m.sN + kRowSlack);
        Sb = dalloc<R>((size_t)5 * m.sC);
        if (mesh_param) { Mb = dalloc<R>((size_t)19 * m.sF); Vb = dalloc<R>((size_t)m.sC); }
        adjoint_ready = true;
    }

    // adjoint step (Function_primal_grad for parameters='source'): W[0] must hold the state at the START of the
    // step; adjoint of the step output in (rhoa, rhoUa, rhoEa); result: adjoint w.r.t. the start state in A[0],
    // source-term gradient accumulated into Sb (static accumulator semantics, adpy/adpy/variable.py:484-490).
    void adjoint_step(R dt, const R* rhoa, const R* rhoUa, const R* rhoEa, R obja) {
        ensure_adjoint_buffers();
        // large single-rank meshes: the adjoint seeds travel host -> device on the side stream while the forward sweep
        // (which only needs the state) already runs; the reverse sweep waits for them
        const bool overlap = !comm && m.nRemoteCells == 0 && m.nInternalCells >= overlap_min_cells;
        if (!overlap) {
            put5(A[3], rhoa, rhoUa, rhoEa);
            adjoint_step_resident(dt, obja, false);
            return;
        }
        ex.side_begin();               // ordered after the state upload issued so far (it uses the same staging buffer)
        put5(A[3], rhoa, rhoUa, rhoEa);
        ex.side_end();
        primal_step(dt, true);
        ex.join();
        adjoint_reverse(dt, obja);
    }
    // same on resident data: adjoint input already in A[3] (chain: take the previous call's result A[0])
    void adjoint_step_resident(R dt, R obja, bool chain) {
        ensure_adjoint_buffers();
        if (chain) { R* t = A[0]; A[0] = A[3]; A[3] = t; }
        with_graph({2ull, epoch, bits((double)dt), bits((double)obja), (unsigned long long)W[0], (unsigned long long)A[0], (unsigned long long)A[3],
                    (unsigned long long)obj.kind, (unsigned long long)obj.patch, (unsigned long long)obj.dir, (unsigned long long)obj.cells,
                    (unsigned long long)Pb, (unsigned long long)(param_patch * 16 + param_key + 1), (unsigned long long)Mb}, [&]() { adjoint_step_body(dt, obja); });
    }
    void adjoint_step_body(R dt, R obja) {
        primal_step(dt, true);
        adjoint_reverse(dt, obja);
    }
    template <int T, int TS> void run_grad_adj_tile(int s, R dt, R obja, int t0, int nt) {
        GradAdjTileBody<R, T, TS> pb;
        pb.ph = ph; pb.m = m; pb.Gb = Gb; pb.Qb = Qb; pb.W = W[s];
        // a_s = sum_{k>=s} alpha[k][s] * a_{k+1}
        pb.A1 = (s <= 0 && RK_ALPHA[0][s] != 0.) ? A[1] : nullptr; pb.c1 = (R)(s <= 0 ? RK_ALPHA[0][s] : 0.);
        pb.A2 = (s <= 1 && RK_ALPHA[1][s] != 0.) ? A[2] : nullptr; pb.c2 = (R)(s <= 1 ? RK_ALPHA[1][s] : 0.);
        pb.A3 = (RK_ALPHA[2][s] != 0.) ? A[3] : nullptr; pb.c3 = (R)RK_ALPHA[2][s];
        pb.objT = (s == 1 && (obj.kind == OBJ_CELL_TV || obj.kind == OBJ_CELL_T)) ? obja : R(0);
        pb.objVol = obj.kind == OBJ_CELL_TV;
        pb.Aout = A[s];
        pb.Sb = (s == 0) ? Sb : nullptr;
        pb.s1 = (R)RK_BETA[0] * dt; pb.s2 = (R)RK_BETA[1] * dt; pb.s3 = (R)RK_BETA[2] * dt;
        run_tiles_range(t0, nt, pb);
    }
    // the reverse flux kernel specialised for the physics of the context (Roe + Sutherland / constant viscosity at compile time)
    template <int T, int TS> void run_flux_grad_tile(int s, R coef, int t0, int nt) {
        if (ph.riemann == RIEMANN_ROE && ph.mu_law == MU_SUTHERLAND)
            run_tiles_range(t0, nt, FluxGradTileBody<R, T, TS, SPEC_ROE_SUTHERLAND>{ph, m, Q[s], G[s], A[s + 1], coef, Qb, Gb});
        else if (ph.riemann == RIEMANN_ROE && ph.mu_law == MU_CONSTANT)
            run_tiles_range(t0, nt, FluxGradTileBody<R, T, TS, SPEC_ROE_CONSTANT>{ph, m, Q[s], G[s], A[s + 1], coef, Qb, Gb});
        else
            run_tiles_range(t0, nt, FluxGradTileBody<R, T, TS, SPEC_GENERIC>{ph, m, Q[s], G[s], A[s + 1], coef, Qb, Gb});
    }
    void adjoint_reverse(R dt, R obja) {
        const int C = m.nInternalCells;
        for (int s = 2; s >= 0; s--) {
            const R coef = (R)(-RK_BETA[s]) * dt;
            const int Ce = early_cells(), Te = early_tiles();
            // late tiles first: they produce the ghost-row adjoints that travel; the early tiles overlap the exchange
            for (int part = 1; part >= 0; part--) {
                const int t0 = part ? Te : 0, nt = part ? m.nTiles - Te : Te;
                if (tile_variant == 0) run_flux_grad_tile<128, 128 + kHalo128r>(s, coef, t0, nt);
                else if (tile_variant == 1) run_flux_grad_tile<128, 128 + kHalo128s>(s, coef, t0, nt);
                else if (tile_variant == 2) run_flux_grad_tile<128, 128 + kHalo128>(s, coef, t0, nt);
                else run_flux_grad_tile<64, 64 + kHalo64>(s, coef, t0, nt);
                if (part) halo_reverse_begin(Gb, 15);
            }
            const R* rG = halo_reverse_end();
            run(nBcells, GhostGradAdjBody<R>{m, bcells, Gb, rG});
            if (s == 1 && obj.kind == OBJ_PLANE_PTLOSS && obja != R(0)) {      // seeds of the cut-plane objective (stage-1 primitives)
                ex.reduce_sum(obj.ncells, PlaneMassBody<R>{ph, m, obj, Q[1]}, red + 4);
                if (comm) comm->allreduce_sum_device(red + 4, 1, ex.stream_handle());
                ex.reduce_sum(obj.ncells, PlaneLossBody<R>{ph, m, obj, Q[1], red + 4}, red + 5); launches += 4;
                if (comm) comm->allreduce_sum_device(red + 5, 1, ex.stream_handle());
                run(obj.ncells, PlaneLossAdjBody<R>{ph, m, obj, Q[1], red + 4, red + 5, obja, Qb});
            }
            if (s == 1 && obj.kind == OBJ_CALLBACK && obja != R(0)) {          // seeds of a traced objective: obja * dJ/d(U,T,p), all rows
                if (!Qseed) Qseed = dalloc<R>((size_t)5 * m.sN);
                ex.zero(Qseed, (size_t)5 * m.sN * sizeof(R));
                obj_fn(obj_user, Q[1], (long long)m.sN, 1, (double)obja, Qseed);
                run(5 * m.sN, AddBody<R>{Qb, Qseed});
            }
            // late tiles first: they complete the ghost rows of U,T,p that travel
            for (int part = 1; part >= 0; part--) {
                const int t0 = part ? Te : 0, nt = part ? m.nTiles - Te : Te;
                if (tile_variant == 0) run_grad_adj_tile<128, 128 + kHalo128r>(s, dt, obja, t0, nt);
                else if (tile_variant == 1) run_grad_adj_tile<128, 128 + kHalo128s>(s, dt, obja, t0, nt);
                else if (tile_variant == 2) run_grad_adj_tile<128, 128 + kHalo128>(s, dt, obja, t0, nt);
                else run_grad_adj_tile<64, 64 + kHalo64>(s, dt, obja, t0, nt);
                if (part) halo_reverse_begin(Qb, 5);
            }
            const R* rQ = halo_reverse_end();
            const R oa = (s == 1) ? obja : R(0);
            run(nBcells, GhostPrimAdjBody<R>{ph, m, obj, oa, bcells, Q[s], Qb, rQ, W[s], A[s]});
            if (param_patch >= 0) run(patches[param_patch].nFaces, BCParamAdjBody<R>{ph, m, obj, oa, param_patch, param_key, Q[s], Qb, Pb});
            if (mesh_param) {
                run(m.nFaces, MeshGradFaceBody<R>{ph, m, obj, oa, coef, Q[s], G[s], A[s + 1], Qb, Gb, f_dunit, f_linw, f_quadw, Mb});
                run(C, MeshGradCellBody<R>{m, Q[s], G[s], Gb, (s == 1 && obj.kind == OBJ_CELL_TV) ? obja : R(0), Vb});
            }
        }
    }
    // ---- mesh metric build on the device (fvm_metrics.h); host AoS in, host AoS out.
    // neighbour: all faces (boundary face b -> ghost cell nIC + b); cellCentres out: [nIC + nBoundary][3]
    void mesh_metrics(int nP, const R* points, int nF, int nIF, int nIC, const int* faces, const int* owner, const int* neighbour,
                      const int* cellFaces, int nPatches, const MetricPatch* mp, const R* remote,
                      R* areas, R* normals, R* faceCentres, R* cellCentres, R* volumes, R* deltas, R* deltasUnit, R* weights,
                      R* linW, R* quadW) {
        const int nB = nF - nIF, nC = nIC + nB;
        if (nP <= 0 || nF <= 0 || nIC <= 0 || nIF < 0 || nIF > nF) throw std::runtime_error("bad mesh sizes");
        for (long i = 0; i < (long)nF * 4; i++) if (faces[i] < 0 || faces[i] >= nP) throw std::runtime_error("face point index out of range");
        for (long f = 0; f < nF; f++) if (owner[f] < 0 || owner[f] >= nIC || neighbour[f] < 0 || neighbour[f] >= nC) throw std::runtime_error("owner/neighbour out of range");
        for (long i = 0; i < (long)nIC * 6; i++) if (cellFaces[i] < 0 || cellFaces[i] >= nF) throw std::runtime_error("cellFaces out of range");
        std::vector<unsigned char> bp(nB + 1, 255);
        if (nPatches > 255) throw std::runtime_error("too many patches");
        for (int p = 0; p < nPatches; p++) {
            if (mp[p].nFaces && (mp[p].startFace < nIF || mp[p].startFace + mp[p].nFaces > nF)) throw std::runtime_error("patch face range out of bounds");
            if (mp[p].kind == 1 && (mp[p].nbrStartFace < nIF || mp[p].nbrStartFace + mp[p].nFaces > nF)) throw std::runtime_error("cyclic partner out of bounds");
            if (mp[p].kind == 2 && mp[p].nFaces && !remote) throw std::runtime_error("processor patch needs the neighbour rank's cell centres");
            for (int i = 0; i < mp[p].nFaces; i++) bp[mp[p].startFace - nIF + i] = (unsigned char)p;
        }
        for (int b = 0; b < nB; b++) if (bp[b] == 255) throw std::runtime_error("boundary face not covered by any patch");
        std::vector<void*> tmp;
        auto up = [&](const void* h, size_t bytes) { void* d = ex.alloc(bytes + 16); tmp.push_back(d); if (h) ex.upload(d, h, bytes); return d; };
        const R* d_pts = (const R*)up(points, (size_t)nP * 3 * sizeof(R));
        const int* d_faces = (const int*)up(faces, (size_t)nF * 4 * 4);
        const int* d_owner = (const int*)up(owner, (size_t)nF * 4);
        const int* d_neigh = (const int*)up(neighbour, (size_t)nF * 4);
        const int* d_cf = (const int*)up(cellFaces, (size_t)nIC * 6 * 4);
        const MetricPatch* d_mp = (const MetricPatch*)up(mp, (size_t)std::max(1, nPatches) * sizeof(MetricPatch));
        const unsigned char* d_bp = (const unsigned char*)up(bp.data(), bp.size());
        const R* d_remote = remote ? (const R*)up(remote, (size_t)nB * 3 * sizeof(R)) : nullptr;
        R* d_n = (R*)up(nullptr, (size_t)nF * 3 * sizeof(R)); R* d_fc = (R*)up(nullptr, (size_t)nF * 3 * sizeof(R));
        R* d_a = (R*)up(nullptr, (size_t)nF * sizeof(R)); R* d_cc = (R*)up(nullptr, (size_t)nC * 3 * sizeof(R));
        R* d_v = (R*)up(nullptr, (size_t)nIC * sizeof(R)); R* d_d = (R*)up(nullptr, (size_t)nF * sizeof(R));
        R* d_du = (R*)up(nullptr, (size_t)nF * 3 * sizeof(R)); R* d_w = (R*)up(nullptr, (size_t)nF * sizeof(R));
        R* d_lw = (R*)up(nullptr, (size_t)nF * 2 * sizeof(R)); R* d_qw = (R*)up(nullptr, (size_t)nF * 6 * sizeof(R));
        run(nF, FaceGeomBody<R>{d_pts, d_faces, d_n, d_fc, d_a});
        run(nIC, CellGeomBody<R>{d_cf, d_n, d_fc, d_a, d_cc, d_v});
        run(nB, GhostCentreBody<R>{d_mp, d_bp, d_owner, d_fc, d_remote, nIF, nIC, d_cc});
        run(nF, FaceWeightBody<R>{d_owner, d_neigh, d_cc, d_fc, d_n, d_d, d_du, d_w, d_lw, d_qw});
        ex.download(areas, d_a, (size_t)nF * sizeof(R)); ex.download(normals, d_n, (size_t)nF * 3 * sizeof(R));
        ex.download(faceCentres, d_fc, (size_t)nF * 3 * sizeof(R)); ex.download(cellCentres, d_cc, (size_t)nC * 3 * sizeof(R));
        ex.download(volumes, d_v, (size_t)nIC * sizeof(R)); ex.download(deltas, d_d, (size_t)nF * sizeof(R));
        ex.download(deltasUnit, d_du, (size_t)nF * 3 * sizeof(R)); ex.download(weights, d_w, (size_t)nF * sizeof(R));
        ex.download(linW, d_lw, (size_t)nF * 2 * sizeof(R)); ex.download(quadW, d_qw, (size_t)nF * 6 * sizeof(R));
        ex.sync();
        for (void* q : tmp) ex.free(q);
    }

    // ---- device-side checkpoint block. The reference's Solver.run(mode='forward') returns every
    // state of a checkpoint block to the host and Adjoint.run feeds them back one by one (adFVM/solver.py:376-382,
    // apps/adjoint.py:217-291); here the block stays in HBM: primal_block stores the state at the start of each of its
    // steps, adjoint_block walks the block backwards with the adjoint fields resident.
    std::vector<R*> block; R* block_red = nullptr; int block_cap = 0;
    void block_reserve(int nsteps) {
        if (nsteps <= (int)block.size() && nsteps <= block_cap) return;
        while ((int)block.size() < nsteps) block.push_back(dalloc<R>((size_t)5 * m.sC + kRowSlack));
        if (nsteps > block_cap) { block_red = dalloc<R>((size_t)2 * nsteps); block_cap = nsteps; }
    }
    // nsteps primal steps from the resident state; dtc[k], objective[k] of step k (rank-local, like get_dtc_obj)
    void primal_block(int nsteps, const double* dt, double* dtc, double* objective) {
        if (!have_mesh || !have_state) throw std::runtime_error("mesh/state not set");
        block_reserve(nsteps);
        const size_t bytes = (size_t)5 * m.sC * sizeof(R);
        for (int k = 0; k < nsteps; k++) {
            ex.copy(block[k], W[0], bytes);
            primal_step((R)dt[k], false);
            ex.copy(block_red + 2 * k, red, 2 * sizeof(R));
        }
        std::vector<R> h((size_t)2 * nsteps);
        ex.download(h.data(), block_red, h.size() * sizeof(R)); ex.sync();
        for (int k = 0; k < nsteps; k++) {
            if (dtc) dtc[k] = (double)h[2 * k];
            if (objective) objective[k] = comm ? comm->allreduce_sum((double)h[2 * k + 1]) : (double)h[2 * k + 1];
        }
    }
    // reverse sweep over the block stored by the last primal_block: the adjoint of the block's final state is resident
    // (A[0], from adjoint_step or a previous adjoint_block); afterwards A[0] holds the adjoint of the block's first state,
    // Sb the accumulated source-term gradient. The resident primal state is left at the block's first state.
    // viscous: the adjoint artificial viscosity is applied after every step (viscousInterval = 1, apps/adjoint.py:250,288-289)
    void adjoint_block(int nsteps, const double* dt, R obja, bool viscous = false) {
        if (nsteps > (int)block.size()) throw std::runtime_error("adjoint_block: no stored primal block of that length");
        ensure_adjoint_buffers();
        const size_t bytes = (size_t)5 * m.sC * sizeof(R);
        for (int k = nsteps - 1; k >= 0; k--) {
            ex.copy(W[0], block[k], bytes);
            adjoint_step_resident((R)dt[k], obja, true);
            if (viscous) adjoint_viscous((R)dt[k]);
        }
    }
    // ---- adjoint artificial viscosity (fvm_viscosity.h): what `primal_grad_viscous` adds to
    // `primal_grad` (apps/adjoint.py:127-141): M_2norm of the step's START state, then one implicit diffusion step of the
    // adjoint fields the reverse sweep produced.
    int visc_type = VISC_NONE, visc_maxit = 500; double visc_scaling = 0., visc_rtol = 0.; long visc_iterations = 0;
    R *vlam = nullptr, *vM = nullptr, *vcf = nullptr, *vdg = nullptr, *vx = nullptr, *vr = nullptr, *vp = nullptr, *vq = nullptr, *vs = nullptr;
    enum { VS_VTOT = 0, VS_N = 1, VS_RZA = 2, VS_RZB = 7, VS_PQ = 12, VS_BZ = 17, VS_SIZE = 24 };
    void set_adjoint_viscosity(int type, double scaling, double rtol, int maxit) {
        if (type < VISC_NONE || type > VISC_UNIFORM) throw std::runtime_error("unknown adjoint viscosity type");
        visc_type = type; visc_scaling = scaling;
        visc_rtol = rtol > 0. ? rtol : (sizeof(R) == 8 ? 1e-13 : 1e-6);
        visc_maxit = maxit > 0 ? maxit : 500;
    }
    void ensure_visc_buffers() {
        if (vlam) return;
        vlam = dalloc<R>((size_t)m.sC); vM = dalloc<R>((size_t)m.sN); vcf = dalloc<R>((size_t)6 * m.sC); vdg = dalloc<R>((size_t)m.sC);
        vx = dalloc<R>((size_t)5 * m.sN); vp = dalloc<R>((size_t)5 * m.sN + kRowSlack /* the tile kernel bulk-copies whole T-cell rows of vp */); vr = dalloc<R>((size_t)5 * m.sC); vq = dalloc<R>((size_t)5 * m.sC);
        vs = dalloc<R>(VS_SIZE);
    }
    // M_2norm (vM [nCells], ghost rows filled with the default boundary + halo) of the state whose primitives (ghost rows
    // filled) and Green-Gauss gradients are Qs / Gs
    void viscosity_field(const R* Qs, const R* Gs) {
        if (visc_type == VISC_NONE) throw std::runtime_error("adjoint viscosity type not set (adfvm_set_adjoint_viscosity)");
        ensure_visc_buffers();
        const int C = m.nInternalCells, nLB = m.nLocalFaces - m.nInternalFaces;
        run(C, ViscEigBody<R>{ph, m, visc_type, Qs, Gs, vlam});
        ex.reduce_sum(C, ViscVolBody<R>{m.vol}, vs + VS_VTOT); ex.reduce_sum(C, ViscNormBody<R>{m.vol, vlam}, vs + VS_N); launches += 4;
        if (comm) comm->allreduce_sum_device(vs + VS_VTOT, 2, ex.stream_handle());
        run(C, ViscScaleBody<R>{vlam, vs, (R)visc_scaling, vM});
        run(nLB, GhostScalarBody<R>{m, vM});
        halo_begin(vM, 1); halo_end();
    }
    // one backward-Euler diffusion step (face conductance dt * areas * DT / deltas) of the volume-weighted adjoint fields Aio [5][sC]
    void viscosity_apply(R* Aio, R dt) {
        const int C = m.nInternalCells;
        run(C, ViscCoefBody<R>{m, vM, dt, vcf, vdg});
        run(C, ViscStartBody<R>{m, Aio, vx});
        halo_begin(vx, 5); halo_end();
        R* rz = vs + VS_RZA; R* rzn = vs + VS_RZB;
        ex.reduce_sum5(C, ViscRhsNormBody<R>{m, vdg, Aio}, vs + VS_BZ);
        ex.reduce_sum5(C, ViscInitBody<R>{m, vcf, vdg, Aio, vx, vr, vp}, rz); launches += 4;
        if (comm) { comm->allreduce_sum_device(vs + VS_BZ, 5, ex.stream_handle()); comm->allreduce_sum_device(rz, 5, ex.stream_handle()); }
        auto converged = [&](const R* z) {
            R h[VS_SIZE]; ex.download(h, vs, sizeof(h)); ex.sync();
            const R* hz = h + (z - vs); bool ok = true;
            for (int k = 0; k < 5; k++) {
                if (!(hz[k] == hz[k]) || !(h[VS_BZ + k] == h[VS_BZ + k])) throw std::runtime_error("adjoint viscosity: NaN in the diffusion solve");
                if (hz[k] < R(0)) throw std::runtime_error("adjoint viscosity: the diffusion operator is indefinite (negative M_2norm?)");
                if ((double)hz[k] > visc_rtol * visc_rtol * (double)h[VS_BZ + k]) ok = false;
            }
            return ok;
        };
        int it = 0;
        bool done = converged(rz);
        while (!done) {
            if (it >= visc_maxit) throw std::runtime_error("adjoint viscosity: the diffusion solve did not converge");
            for (int inner = 0; inner < 4; inner++, it++) {            // host looks at the residual every fourth iteration
                halo_begin(vp, 5); halo_end();
                visc_spmv();
                if (comm) comm->allreduce_sum_device(vs + VS_PQ, 5, ex.stream_handle());
                if (m.T == 128) run_tiles_range(0, m.nTiles, ViscUpdateTileBody<R, 128>{m, vdg, vp, vq, rz, vs + VS_PQ, vx, vr, vpart});
                else run_tiles_range(0, m.nTiles, ViscUpdateTileBody<R, 64>{m, vdg, vp, vq, rz, vs + VS_PQ, vx, vr, vpart});
                ex.reduce_sum5(m.nTiles, ViscPartialBody<R>{vpart}, rzn); launches += 4;
                if (comm) comm->allreduce_sum_device(rzn, 5, ex.stream_handle());
                run(C, ViscDirBody<R>{m, vdg, vr, rz, rzn, vp});
                std::swap(rz, rzn);
            }
            done = converged(rz);
        }
        visc_iterations = it;
        run(C, ViscFinishBody<R>{m, vx, Aio});
    }
    // q = A p and the five dot products p . q: tile kernel (search direction staged in shared memory) + sum of the per-tile partials
    R* vpart = nullptr;
    template <int T, int TS> void run_visc_spmv_tile() { run_tiles_range(0, m.nTiles, ViscSpmvTileBody<R, T, TS>{m, vcf, vdg, vp, vq, vpart}); }
    void visc_spmv() {
        if (!vpart) vpart = dalloc<R>((size_t)5 * m.nTiles + 8);
        if (tile_variant == 0) run_visc_spmv_tile<128, 128 + kHalo128r>();
        else if (tile_variant == 1) run_visc_spmv_tile<128, 128 + kHalo128s>();
        else if (tile_variant == 2) run_visc_spmv_tile<128, 128 + kHalo128>();
        else run_visc_spmv_tile<64, 64 + kHalo64>();
        ex.reduce_sum5(m.nTiles, ViscPartialBody<R>{vpart}, vs + VS_PQ);
    }
    void adjoint_viscous(R dt) { viscosity_field(Q[0], G[0]); viscosity_apply(A[0], dt); }
    // diagnostic (the reference's write_M_2norm, apps/adjoint.py:26,139): M_2norm [nCells][1] of a host state, reference numbering
    void get_adjoint_viscosity(const R* rho, const R* rhoU, const R* rhoE, R* M_out) {
        if (!have_mesh) throw std::runtime_error("mesh not set");
        check_bcs();
        adjoint_state_begin(rho, rhoU, rhoE);
        try { stage(0, R(0), Q[0], G[0], nullptr, false, false); viscosity_field(Q[0], G[0]); }
        catch (...) { adjoint_state_end(); throw; }
        adjoint_state_end();
        const int C = m.nInternalCells, N = m.nCells;
        run(C, SoaToAosBody<R>{vM, stage_aos, 1, m.sC, m.cell_perm});
        ex.download(M_out, stage_aos, (size_t)C * sizeof(R));
        ex.download(M_out + C, vM + C, (size_t)(N - C) * sizeof(R));
        ex.sync();
    }
    void put_adjoint(const R* rhoa, const R* rhoUa, const R* rhoEa) { ensure_adjoint_buffers(); put5(A[0], rhoa, rhoUa, rhoEa); }
    void get_adjoint(R* rhoa, R* rhoUa, R* rhoEa) {
        if (!adjoint_ready) throw std::runtime_error("no adjoint data: call primal_grad / set_adjoint first");
        get5(A[0], rhoa, rhoUa, rhoEa);
    }
    void get_source_grad(R* a, R* b, R* c, bool zero_after) {
        if (!adjoint_ready) throw std::runtime_error("no adjoint data: call primal_grad / set_adjoint first");
        get5(Sb, a, b, c);
        if (zero_after) ex.zero(Sb, (size_t)5 * m.sC * sizeof(R));
    }
};

}  // namespace fvm
