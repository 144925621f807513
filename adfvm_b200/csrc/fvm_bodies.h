// Per-element bodies (functors) of every kernel on the residual path and of its reverse sweep.
// A body is a plain struct of pointers + an `operator()(int i)`; the CUDA side wraps it in a
// __global__ grid-stride launcher (fvm_cuda_exec.cuh), the CPU test simulator in a for loop.
//
// Device data layout (all SoA, component-major, 32-element padded strides):
//   W  [5][sC]   conserved state (rho, rhoU_x, rhoU_y, rhoU_z, rhoE), internal cells
//   Q  [5][sN]   primitive state (U_x, U_y, U_z, T, p), internal + ghost cells
//   G  [15][sN]  gradients: dU_i/dx_j at 3*i+j, dT/dx_j at 9+j, dp/dx_j at 12+j
//   face metrics [k][sF]; cell connectivity [6][sC]
// No scatter of the reference (Tensor.collate -> atomicAdd, adpy/adpy/tensor.py:393-394) is an atomic here: the
// flux kernels (fvm_tile_bodies.h) sum the face contributions of a cell in the registers of the lane that owns it, in
// the fixed order of the sub-tile's schedule (fvm_tiles.h); the gradient kernels are cell-centred gathers over the six
// faces of the hexahedron. All sums have a fixed order: results are bitwise reproducible run to run, primal and adjoint.
#pragma once
#include "fvm_math.h"

namespace fvm {

enum BCType { BC_CALCULATED = 0, BC_CYCLIC = 1, BC_ZEROGRADIENT = 2, BC_FIXEDVALUE = 3, BC_SYMMETRY = 4,
              BC_CBC_UPT = 5, BC_CBC_TOTAL_PT = 6, BC_PROCESSOR = 7 };
enum ObjKind { OBJ_NONE = 0, OBJ_CELL_TV = 1, OBJ_PATCH_PA = 2, OBJ_DRAG = 3, OBJ_PLANE_PTLOSS = 4,
               OBJ_CELL_T = 5,       // sum of T over the cells, no volume weight (reference templates/box.py:10-16)
               OBJ_CALLBACK = 6 };   // evaluated by the host layer from the case file's traced kernels (Solver::obj_fn)
enum { MAX_PATCHES = 255 };

template <typename R> struct PatchDev {
    int startFace, nFaces, cellStartFace;
    int kind;                 // FaceKind used by the flux
    int bc[3];                // BCType of U, T, p
    int gbc;                  // BC_CYCLIC or BC_ZEROGRADIENT (gradient fields) / BC_PROCESSOR
    int nbrStartFace, nbrCellStartFace;   // cyclic partner
    // SoA value arrays [d][nFaces] (NULL when the BC has no such input, adFVM/BCs.py createInput)
    const R *valU, *valT, *valp, *U0, *T0, *p0, *Tt, *pt, *dir;
};

template <typename R> struct MeshDev {
    int nCells, nFaces, nInternalCells, nInternalFaces, nLocalCells, nRemoteCells, nLocalFaces, nGhostCells;
    int sC, sN, sF;
    const R *area, *normal, *weight, *idelta, *vol;   // face-indexed SoA (idelta = 1/deltas); the other flux metrics live in the chunks
    const R *cfm;                        // cell-face metrics of the Green-Gauss gradient [24][sC] (CellFaceMetricBody)
    const int *owner, *neigh, *cellFaces, *cellNbr;
    const unsigned char *cellOwner;      // bit j set: the cell owns its j-th face
    const unsigned char *bpatch;         // [nGhostCells] patch index of each boundary face
    const PatchDev<R>* patches;
    int nPatches;
    // tiles (fvm_tiles.h): tile t owns cells [t*T, min(C,(t+1)*T)); its sub-tile w (32 cells, one warp) owns the per-round
    // chunks [round_start[t*T/32+w], round_start[t*T/32+w+1])
    int T, nTiles;
    const int* round_start; const R* chunks;
    const int* halo_round;               // [nTiles*T/32] first round (relative to round_start) whose entries read halo slots
    const int* halo_start; const int* halo_cell;    // tile t's halo slots T.. hold cells halo_cell[halo_start[t]..halo_start[t+1])
    const int* cell_perm;                // [C] device cell -> reference (host) cell
    const unsigned short* nbrSlot;       // [6][sC] tile slot of the cell's j-th neighbour (own cells [0,T), halo slots from T on); bit 15: ghost cell
};

// OBJ_PLANE_PTLOSS (reference adFVM/objectives/vane.py:36-66): cells / areas of a cut plane (device cell numbering),
// inlet total pressure ptin, plane normal nrm, factor scale (the a = 0.4 of vane.py:120)
template <typename R> struct ObjDev { int kind, patch, dir; const int* cells; const R* areas; int ncells; R ptin, scale, nrm[3]; };

// ------------------------------------------------------------------------------------------ loads
template <typename R> FVM_HD void load_prim(const R* Q, int sN, int c, Prim<R>& q) {
    q.U[0] = Q[c]; q.U[1] = Q[sN + c]; q.U[2] = Q[2 * sN + c]; q.T = Q[3 * sN + c]; q.p = Q[4 * sN + c];
}
template <typename R> FVM_HD void store_prim(R* Q, int sN, int c, const Prim<R>& q) {
    Q[c] = q.U[0]; Q[sN + c] = q.U[1]; Q[2 * sN + c] = q.U[2]; Q[3 * sN + c] = q.T; Q[4 * sN + c] = q.p;
}
template <typename R> FVM_HD void add_prim(R* Q, int sN, int c, const Prim<R>& q) {
    Q[c] += q.U[0]; Q[sN + c] += q.U[1]; Q[2 * sN + c] += q.U[2]; Q[3 * sN + c] += q.T; Q[4 * sN + c] += q.p;
}
template <typename R> FVM_HD void load_grad(const R* G, int sN, int c, Grad<R>& g) {
    for (int k = 0; k < 9; k++) g.U[k] = G[k * sN + c];
    for (int k = 0; k < 3; k++) { g.T[k] = G[(9 + k) * sN + c]; g.p[k] = G[(12 + k) * sN + c]; }
}
template <typename R> FVM_HD void store_grad(R* G, int sN, int c, const Grad<R>& g) {
    for (int k = 0; k < 9; k++) G[k * sN + c] = g.U[k];
    for (int k = 0; k < 3; k++) { G[(9 + k) * sN + c] = g.T[k]; G[(12 + k) * sN + c] = g.p[k]; }
}
template <typename R> FVM_HD int face_kind(const MeshDev<R>& m, int f) {
    return f < m.nInternalFaces ? (int)FACE_COUPLED : m.patches[m.bpatch[f - m.nInternalFaces]].kind;
}

// halo buffers are patch-major: for each processor patch (in face order) a block [ncomp][nFaces_p], so that one
// patch is one contiguous message. Offset of component 0 of remote face r; component stride returned in cs.
template <typename R> FVM_HD long halo_offset(const MeshDev<R>& m, int r, int ncomp, int& cs) {
    const int f = m.nLocalFaces + r;
    const PatchDev<R>& P = m.patches[m.bpatch[f - m.nInternalFaces]];
    cs = P.nFaces;
    return (long)(P.startFace - m.nLocalFaces) * ncomp + (f - P.startFace);
}

// ------------------------------------------------------------------------------------------ a1
// conserved -> primitive on internal cells (adFVM/density.py:162-171)
template <typename R> struct PrimitiveBody {
    static constexpr const char* kName = "primitive";
    Phys<R> ph; int sC, sN; const R* W; R* Q;
    FVM_HD void operator()(int c) const {
        R rhoU[3] = {W[sC + c], W[2 * sC + c], W[3 * sC + c]};
        Prim<R> q; primitive(ph, W[c], rhoU, W[4 * sC + c], q);
        store_prim(Q, sN, c, q);
    }
};

// ------------------------------------------------------------------------------------------ a2/a3
// Ghost rows of U,T,p for all LOCAL boundary faces in one launch. Every BC reads internal cells only and
// writes its own ghost row, so the patch order of the reference (sorted names, then characteristic,
// adFVM/field.py:148-156, adFVM/density.py:333-337,416-418) does not affect the result.
template <typename R> struct GhostPrimBody {
    static constexpr const char* kName = "ghost_prim";
    Phys<R> ph; MeshDev<R> m; R* Q;
    FVM_HD void operator()(int b) const {
        const int f = m.nInternalFaces + b, g = m.nInternalCells + b;
        const PatchDev<R>& P = m.patches[m.bpatch[b]];
        const int i = f - P.startFace, nf = P.nFaces, sN = m.sN;
        const int own = m.owner[f];
        // U
        switch (P.bc[0]) {
        case BC_CYCLIC: { int s = m.owner[P.nbrStartFace + i]; for (int k = 0; k < 3; k++) Q[k * sN + g] = Q[k * sN + s]; } break;
        case BC_ZEROGRADIENT: for (int k = 0; k < 3; k++) Q[k * sN + g] = Q[k * sN + own]; break;
        case BC_FIXEDVALUE: for (int k = 0; k < 3; k++) Q[k * sN + g] = P.valU[k * nf + i]; break;
        case BC_SYMMETRY: {
            R u[3] = {Q[own], Q[sN + own], Q[2 * sN + own]};
            R n[3] = {m.normal[f], m.normal[m.sF + f], m.normal[2 * m.sF + f]};
            R un = dot3(u, n);
            for (int k = 0; k < 3; k++) Q[k * sN + g] = u[k] - un * n[k];
        } break;
        default: break;
        }
        // T, p
        for (int fld = 1; fld < 3; fld++) {
            const int k = 2 + fld;
            switch (P.bc[fld]) {
            case BC_CYCLIC: Q[k * sN + g] = Q[k * sN + m.owner[P.nbrStartFace + i]]; break;
            case BC_ZEROGRADIENT: case BC_SYMMETRY: Q[k * sN + g] = Q[k * sN + own]; break;
            case BC_FIXEDVALUE: Q[k * sN + g] = (fld == 1 ? P.valT : P.valp)[i]; break;
            default: break;
            }
        }
        // characteristic BCs hang off the p field (adFVM/BCs.py:139-184)
        if (P.bc[2] == BC_CBC_UPT) {
            for (int k = 0; k < 3; k++) Q[k * sN + g] = P.U0[k * nf + i];
            Q[3 * sN + g] = P.T0[i]; Q[4 * sN + g] = P.p0[i];
        } else if (P.bc[2] == BC_CBC_TOTAL_PT) {
            R d[3];
            for (int k = 0; k < 3; k++) d[k] = P.dir ? P.dir[k * nf + i] : m.normal[k * m.sF + f];
            R u[3] = {Q[own], Q[sN + own], Q[2 * sN + own]};
            R To = Q[3 * sN + own];
            R Un = dot3(u, d);
            R Tt = P.Tt[i];
            for (int k = 0; k < 3; k++) Q[k * sN + g] = Un * d[k];
            Q[3 * sN + g] = Tt - R(0.5) * Un * Un / ph.Cp;
            Q[4 * sN + g] = P.pt[i] * pow(To / Tt, ph.gamma / (ph.gamma - R(1)));
        }
    }
};

// ghost rows of gradU, gradT, gradp: cyclic copies the partner's owner row, everything else the own
// owner row (gradient fields carry the mesh default boundary: adFVM/density.py:133, mesh.py:702-711)
template <typename R> struct GhostGradBody {
    static constexpr const char* kName = "ghost_grad";
    MeshDev<R> m; R* G;
    FVM_HD void operator()(int b) const {
        const int f = m.nInternalFaces + b, g = m.nInternalCells + b;
        const PatchDev<R>& P = m.patches[m.bpatch[b]];
        int s = (P.gbc == BC_CYCLIC) ? m.owner[P.nbrStartFace + (f - P.startFace)] : m.owner[f];
        for (int k = 0; k < 15; k++) G[k * m.sN + g] = G[k * m.sN + s];
    }
};

// ------------------------------------------------------------------------------------------ a4
// Green-Gauss cell gradient (adFVM/op.py:45-63)
template <typename R> struct GradCellBody {
    static constexpr const char* kName = "grad_cell";
    MeshDev<R> m; const R* Q; R* G;
    FVM_HD void operator()(int c) const {
        Prim<R> qc; load_prim(Q, m.sN, c, qc);
        Grad<R> g; zero(g);
        const unsigned ob = m.cellOwner[c];
        for (int j = 0; j < 6; j++) {
            const int f = m.cellFaces[j * m.sC + c], nb = m.cellNbr[j * m.sC + c];
            const bool own = (ob >> j) & 1u;
            const R S = m.area[f], w = m.weight[f];
            const R sg = own ? S : -S;
            const R SN[3] = {sg * m.normal[f], sg * m.normal[m.sF + f], sg * m.normal[2 * m.sF + f]};
            const R wp = own ? R(1) - w : w;
            Prim<R> qn; load_prim(Q, m.sN, nb, qn);
            const R a = R(1) - wp;
            for (int i = 0; i < 3; i++) {
                R pf = qc.U[i] * a + qn.U[i] * wp;
                for (int k = 0; k < 3; k++) g.U[3 * i + k] += pf * SN[k];
            }
            R tf = qc.T * a + qn.T * wp, pf = qc.p * a + qn.p * wp;
            for (int k = 0; k < 3; k++) { g.T[k] += tf * SN[k]; g.p[k] += pf * SN[k]; }
        }
        const R iv = R(1) / m.vol[c];
        for (int k = 0; k < 9; k++) g.U[k] = g.U[k] * iv;       // reference divides (gradPhi/volumes)
        for (int k = 0; k < 3; k++) { g.T[k] = g.T[k] * iv; g.p[k] = g.p[k] * iv; }
        store_grad(G, m.sN, c, g);
    }
};

// ------------------------------------------------------------------------------------------ a13
// objective contributions; reduced with a fixed-order tree
template <typename R> struct ObjectiveBody {
    static constexpr const char* kName = "objective";
    Phys<R> ph; MeshDev<R> m; ObjDev<R> o; const R* Q;
    FVM_HD R operator()(int i) const {
        if (o.kind == OBJ_CELL_TV) return Q[3 * m.sN + i] * m.vol[i];
        if (o.kind == OBJ_CELL_T) return Q[3 * m.sN + i];
        const PatchDev<R>& P = m.patches[o.patch];
        const int f = P.startFace + i, g = m.neigh[f];
        if (o.kind == OBJ_PATCH_PA) return Q[4 * m.sN + g] * m.area[f];
        // drag (reference templates/cylinder_test.py:9-19)
        const int own = m.owner[f];
        R mu = viscosity(ph, Q[3 * m.sN + g]);
        R mung = mu * (Q[o.dir * m.sN + g] - Q[o.dir * m.sN + own]) * m.idelta[f];
        return (Q[4 * m.sN + g] * m.normal[o.dir * m.sF + f] - mung) * m.area[f];
    }
};
// ---- mass-flow averaged total-pressure loss over a cut plane: obj = scale * S2/S1, S1 = sum m_i,
// S2 = sum l_i m_i, m_i = rho_i (U_i.n) A_i, l_i = (ptin - pt_i)/ptin, pt = p (1 + (g-1)/2 M^2)^(g/(g-1))
template <typename R> FVM_HD void plane_cell(const Phys<R>& ph, const ObjDev<R>& o, const R* Q, int sN, int i, R& mflux, R& loss, Prim<R>& dm, Prim<R>& dl) {
    const int c = o.cells[i];
    Prim<R> q; load_prim(Q, sN, c, q);
    const R A = o.areas[i], g = ph.gamma, gm1 = ph.gm1;
    const R Rg = ph.Cv * gm1;                       // rho = p / (Cv T (g-1))
    const R rho = q.p / (Rg * q.T);
    const R un = dot3(q.U, o.nrm);
    mflux = rho * un * A;
    const R c2 = g * q.p / rho;
    const R M2 = dot3(q.U, q.U) / c2;
    const R B = R(1) + R(0.5) * gm1 * M2, e = ph.g_gm1;
    const R Be1 = pow(B, e - R(1));
    const R pt = q.p * Be1 * B;
    loss = (o.ptin - pt) / o.ptin;
    // derivatives w.r.t. (U, T, p)
    for (int k = 0; k < 3; k++) dm.U[k] = rho * o.nrm[k] * A;
    dm.p = mflux / q.p; dm.T = -mflux / q.T;
    const R f = q.p * e * Be1 * R(0.5) * gm1;       // d pt / d M2
    const R ip = -R(1) / o.ptin;
    for (int k = 0; k < 3; k++) dl.U[k] = ip * f * R(2) * q.U[k] / c2;
    dl.T = ip * f * (-M2 / q.T);                    // c2 = g (Cv (g-1)) T
    dl.p = ip * Be1 * B;
}
template <typename R> struct PlaneMassBody {       // pass 1: S1
    static constexpr const char* kName = "objective";
    Phys<R> ph; MeshDev<R> m; ObjDev<R> o; const R* Q;
    FVM_HD R operator()(int i) const { R mf, l; Prim<R> a, b; plane_cell(ph, o, Q, m.sN, i, mf, l, a, b); return mf; }
};
template <typename R> struct PlaneLossBody {       // pass 2: scale * sum l_i m_i / S1  (S1 read from device memory)
    static constexpr const char* kName = "objective";
    Phys<R> ph; MeshDev<R> m; ObjDev<R> o; const R* Q; const R* S1;
    FVM_HD R operator()(int i) const { R mf, l; Prim<R> a, b; plane_cell(ph, o, Q, m.sN, i, mf, l, a, b); return o.scale * l * mf / S1[0]; }
};
// reverse: Qb[cell] += obja * d obj / d(U,T,p); S1 = sum m, P2 = obj (= scale * S2/S1) from the two passes above
template <typename R> struct PlaneLossAdjBody {
    static constexpr const char* kName = "objective_adj";
    Phys<R> ph; MeshDev<R> m; ObjDev<R> o; const R* Q; const R* S1; const R* P2; R obja; R* Qb;
    FVM_HD void operator()(int i) const {
        R mf, l; Prim<R> dm, dl; plane_cell(ph, o, Q, m.sN, i, mf, l, dm, dl);
        const R s = obja * o.scale / S1[0];
        const R lbar = P2[0] / o.scale;             // S2/S1
        Prim<R> qb;
        for (int k = 0; k < 3; k++) qb.U[k] = s * ((l - lbar) * dm.U[k] + mf * dl.U[k]);
        qb.T = s * ((l - lbar) * dm.T + mf * dl.T);
        qb.p = s * ((l - lbar) * dm.p + mf * dl.p);
        add_prim(Qb, m.sN, o.cells[i], qb);          // the cells of a cut plane are distinct
    }
};
// host cell ids -> device cell ids: inverse of cell_perm, then gather
template <typename R> struct InvertPermBody {
    static constexpr const char* kName = "invert_perm";
    const int* perm; int* inv;
    FVM_HD void operator()(int i) const { inv[perm[i]] = i; }
};
template <typename R> struct GatherIntBody {
    static constexpr const char* kName = "gather_int";
    const int* map; const int* in; int* out;
    FVM_HD void operator()(int i) const { out[i] = map[in[i]]; }
};

// adjoint of the objective w.r.t. the ghost row of boundary face f (scaled by obja) ...
template <typename R> FVM_HD void objective_ghost_adj(const Phys<R>& ph, const MeshDev<R>& m, const ObjDev<R>& o, const R* Q,
                                                      R obja, int f, Prim<R>& qb) {
    if (o.kind != OBJ_PATCH_PA && o.kind != OBJ_DRAG) return;
    const PatchDev<R>& P = m.patches[o.patch];
    if (f < P.startFace || f >= P.startFace + P.nFaces) return;
    const int g = m.neigh[f];
    if (o.kind == OBJ_PATCH_PA) { qb.p += obja * m.area[f]; return; }
    const int own = m.owner[f];
    R Tg = Q[3 * m.sN + g];
    R mu = viscosity(ph, Tg);
    R du = Q[o.dir * m.sN + g] - Q[o.dir * m.sN + own];
    R A = m.area[f], idel = m.idelta[f];
    qb.p += obja * m.normal[o.dir * m.sF + f] * A;
    qb.U[o.dir] += -obja * mu * idel * A;
    qb.T += -obja * viscosity_dT(ph, Tg, mu) * du * idel * A;
}
// ... and w.r.t. the owner cell of boundary face f
template <typename R> FVM_HD void objective_owner_adj(const Phys<R>& ph, const MeshDev<R>& m, const ObjDev<R>& o, const R* Q,
                                                      R obja, int f, Prim<R>& qb) {
    if (o.kind != OBJ_DRAG) return;
    const PatchDev<R>& P = m.patches[o.patch];
    if (f < P.startFace || f >= P.startFace + P.nFaces) return;
    R mu = viscosity(ph, Q[3 * m.sN + m.neigh[f]]);
    qb.U[o.dir] += obja * mu * m.area[f] * m.idelta[f];
}

// ========================================================================================== reverse
// A. adjoint of the flux + scatter (+ RK residual weight): FluxGradTileBody in fvm_tile_bodies.h

// B. adjoint of the gradient ghost fill, gathered per boundary-adjacent cell (list bcells):
// Gb[c] += Gb[ghost rows that copied from c]. Processor faces add the rows received from the peer.
template <typename R> struct GhostGradAdjBody {
    static constexpr const char* kName = "ghost_grad_adj";
    MeshDev<R> m; const int* bcells; R* Gb; const R* recvG;   // recvG: patch-major halo buffer or NULL
    FVM_HD void operator()(int i) const {
        const int c = bcells[i];
        R acc[15];
        for (int k = 0; k < 15; k++) acc[k] = R(0);
        bool any = false;
        for (int j = 0; j < 6; j++) {
            const int f = m.cellFaces[j * m.sC + c];
            if (f < m.nInternalFaces) continue;
            if (f >= m.nLocalFaces) {
                if (recvG) { int cs; const long o = halo_offset(m, f - m.nLocalFaces, 15, cs); for (int k = 0; k < 15; k++) acc[k] += recvG[o + (long)k * cs]; any = true; }
                continue;
            }
            const PatchDev<R>& P = m.patches[m.bpatch[f - m.nInternalFaces]];
            const int g = (P.gbc == BC_CYCLIC) ? P.nbrCellStartFace + (f - P.startFace) : m.nInternalCells + (f - m.nInternalFaces);
            for (int k = 0; k < 15; k++) acc[k] += Gb[k * m.sN + g];
            any = true;
        }
        // internal rows of Gb hold the gradient adjoint already divided by the cell volume (FluxGradTileBody::finish)
        if (any) { const R iv = rcp(m.vol[c]); for (int k = 0; k < 15; k++) Gb[k * m.sN + c] += acc[k] * iv; }
    }
};

// C + F. adjoint of gradCell fused with the adjoint of primitive() and of the RK combination, one thread per cell.
//   Green-Gauss (adFVM/op.py:45-63): G_c = 1/V_c sum_j (a_j phi_c + (1-a_j) phi_nb_j) SN_j, with SN_j = +-S_f n_f
//   outward of c and a_j the weight of the cell's own value (cfm rows 4j..4j+3). With H = Gb/V (internal rows of Gb
//   are stored that way), the face shared by c and nb appears in G_c with weight a_j, normal SN_j and in G_nb with
//   weight a_j on phi_c and normal -SN_j:   Qb_c += sum_j a_j SN_j . (H_c - H_nb_j);
//   a ghost neighbour has no gradient of its own and receives (1-a_j) SN_j . H_c in its Qb row (exclusive writer).
//   Then a_s = sum_k alpha_k a_k + (dQ/dW)^T Qb_c (+ cell objective seed, + source-gradient accumulation on the last
//   reverse stage). Cells next to a boundary get the share of their ghost rows later (GhostPrimAdjBody, linear).
// (the kernel itself is GradAdjTileBody in fvm_tile_bodies.h: one CTA per tile, the gradient adjoints of the tile and of its halo
// staged in shared memory)
// cell-face metrics (mesh upload, one-off): rows 4j..4j+2 = +-S_f n_f (outward of the cell), row 4j+3 = weight of the
// cell's own value in the face interpolate (adFVM/op.py:52-58: w' = w + o - 2 w o is the NEIGHBOUR's weight)
template <typename R> struct CellFaceMetricBody {
    static constexpr const char* kName = "cell_face_metric";
    MeshDev<R> m; R* cfm;
    FVM_HD void operator()(int c) const {
        const unsigned ob = m.cellOwner[c];
        for (int j = 0; j < 6; j++) {
            const int f = m.cellFaces[j * m.sC + c];
            const bool own = (ob >> j) & 1u;
            const R S = m.area[f], w = m.weight[f];
            const R sg = own ? S : -S;
            for (int k = 0; k < 3; k++) cfm[(long)(4 * j + k) * m.sC + c] = sg * m.normal[(long)k * m.sF + f];
            const R wp = own ? R(1) - w : w;
            cfm[(long)(4 * j + 3) * m.sC + c] = R(1) - wp;
        }
    }
};

// E. adjoint of the U,T,p ghost fill (+ objective seeds on boundary faces), gathered per boundary-adjacent cell.
template <typename R> struct GhostPrimAdjBody {
    static constexpr const char* kName = "ghost_prim_adj";
    Phys<R> ph; MeshDev<R> m; ObjDev<R> o; R obja;   // obja == 0 on stages that do not carry the objective
    const int* bcells; const R* Q; const R* Qb; const R* recvQ;   // recvQ: patch-major halo buffer or NULL
    const R* W; R* Aout;       // the ghost rows' share goes straight into the stage adjoint: Aout[c] += (dQ/dW)^T acc (linear)
    FVM_HD void ghost_total(int f, Prim<R>& q) const {
        load_prim(Qb, m.sN, m.nInternalCells + (f - m.nInternalFaces), q);
        if (obja != R(0)) objective_ghost_adj(ph, m, o, Q, obja, f, q);
    }
    FVM_HD void operator()(int i) const {
        const int c = bcells[i];
        Prim<R> acc; zero(acc);
        for (int j = 0; j < 6; j++) {
            const int f = m.cellFaces[j * m.sC + c];
            if (f < m.nInternalFaces) continue;
            if (f >= m.nLocalFaces) {
                if (recvQ) { int n; const long r = halo_offset(m, f - m.nLocalFaces, 5, n);
                    acc.U[0] += recvQ[r]; acc.U[1] += recvQ[n + r]; acc.U[2] += recvQ[2 * (long)n + r];
                    acc.T += recvQ[3 * (long)n + r]; acc.p += recvQ[4 * (long)n + r]; }
                continue;
            }
            const PatchDev<R>& P = m.patches[m.bpatch[f - m.nInternalFaces]];
            const int idx = f - P.startFace;
            Prim<R> own, par; bool have_own = false, have_par = false;
            // U
            if (P.bc[0] == BC_CYCLIC) { ghost_total(P.nbrStartFace + idx, par); have_par = true; for (int k = 0; k < 3; k++) acc.U[k] += par.U[k]; }
            else if (P.bc[0] == BC_ZEROGRADIENT) { ghost_total(f, own); have_own = true; for (int k = 0; k < 3; k++) acc.U[k] += own.U[k]; }
            else if (P.bc[0] == BC_SYMMETRY) {
                ghost_total(f, own); have_own = true;
                R n[3] = {m.normal[f], m.normal[m.sF + f], m.normal[2 * m.sF + f]};
                R un = dot3(own.U, n);
                for (int k = 0; k < 3; k++) acc.U[k] += own.U[k] - un * n[k];
            }
            // T, p
            for (int fld = 1; fld < 3; fld++) {
                const int bc = P.bc[fld];
                R v = R(0);
                if (bc == BC_CYCLIC) { if (!have_par) { ghost_total(P.nbrStartFace + idx, par); have_par = true; } v = (fld == 1) ? par.T : par.p; }
                else if (bc == BC_ZEROGRADIENT || bc == BC_SYMMETRY) { if (!have_own) { ghost_total(f, own); have_own = true; } v = (fld == 1) ? own.T : own.p; }
                if (fld == 1) acc.T += v; else acc.p += v;
            }
            if (P.bc[2] == BC_CBC_TOTAL_PT) {              // adFVM/BCs.py:178-184
                if (!have_own) { ghost_total(f, own); have_own = true; }
                R d[3];
                for (int k = 0; k < 3; k++) d[k] = P.dir ? P.dir[k * P.nFaces + idx] : m.normal[k * m.sF + f];
                R u[3] = {Q[c], Q[m.sN + c], Q[2 * m.sN + c]};
                R To = Q[3 * m.sN + c], Tt = P.Tt[idx];
                R Un = dot3(u, d);
                R Unb = dot3(own.U, d) - own.T * Un / ph.Cp;
                for (int k = 0; k < 3; k++) acc.U[k] += Unb * d[k];
                R ex = ph.gamma / (ph.gamma - R(1));
                acc.T += own.p * P.pt[idx] * ex * pow(To / Tt, ex - R(1)) / Tt;
            }
            if (obja != R(0)) objective_owner_adj(ph, m, o, Q, obja, f, acc);
        }
        const int sC = m.sC;
        R rhoU[3] = {W[sC + c], W[2 * sC + c], W[3 * sC + c]};
        R out[5] = {0, 0, 0, 0, 0};
        primitive_vjp(ph, W[c], rhoU, W[4 * sC + c], acc, out[0], out + 1, out[4]);
        for (int k = 0; k < 5; k++) Aout[k * sC + c] += out[k];
    }
};

// G. gradient with respect to ONE boundary-condition input array (parameters = ('BCs', field, patch, key), reference
// apps/adjoint.py:108-116): the adjoint of the patch's ghost rows (+ objective seeds) pushed through the BC formula
// (adFVM/BCs.py:111-184), accumulated over stages and steps into Pb [d][nFaces] like the source-term gradient.
template <typename R> struct BCParamAdjBody {
    static constexpr const char* kName = "bc_param_adj";
    Phys<R> ph; MeshDev<R> m; ObjDev<R> o; R obja; int patch, key;   // key: Solver::BCKey
    const R* Q; const R* Qb; R* Pb;
    FVM_HD void operator()(int i) const {
        const PatchDev<R>& P = m.patches[patch];
        const int f = P.startFace + i, nf = P.nFaces, gcell = m.nInternalCells + (f - m.nInternalFaces);
        Prim<R> q; load_prim(Qb, m.sN, gcell, q);
        if (obja != R(0)) objective_ghost_adj(ph, m, o, Q, obja, f, q);
        switch (key) {
        case 0: case 3: for (int k = 0; k < 3; k++) Pb[k * nf + i] += q.U[k]; break;      // U value / U0: ghost U = input
        case 1: case 4: Pb[i] += q.T; break;                                              // T value / T0
        case 2: case 5: Pb[i] += q.p; break;                                              // p value / p0
        case 6: case 7: {                                                                 // Tt, pt of CBC_TOTAL_PT (BCs.py:178-184)
            const int own = m.owner[f];
            const R To = Q[3 * m.sN + own], Tt = P.Tt[i], ex = ph.g_gm1;
            const R ratio = pow(To / Tt, ex);
            if (key == 7) Pb[i] += q.p * ratio;                                           // p_b = pt (To/Tt)^ex
            else Pb[i] += q.T - q.p * P.pt[i] * ratio * ex / Tt;                           // T_b = Tt - Un^2/(2 Cp)
        } break;
        default: break;
        }
    }
};

// H. mesh sensitivities (parameters = 'mesh', reference apps/adjoint.py:105-107): gradient with respect to the ten metric
// arrays of adFVM/mesh.py:27-31. One thread per face, after the reverse sweep of a stage (needs abar, the complete
// gradient adjoints H = Gb/V, the ghost-row adjoints Qb and the forward Q, G of the stage); every output row belongs to
// one face or one cell: no scatter. Accumulated over stages and steps like the source-term gradient.
//   Mb rows [19][sF]: 0 areas, 1 volumesL, 2 volumesR, 3 weights, 4 deltas, 5-7 normals, 8-10 deltasUnit, 11-12 linearWeights,
//   13-18 quadraticWeights (device face numbering)
template <typename R> struct MeshGradFaceBody {
    static constexpr const char* kName = "mesh_grad_face";
    Phys<R> ph; MeshDev<R> m; ObjDev<R> o; R obja; R coef;
    const R *Q, *G, *abar, *Qb, *Gb;              // Gb: internal rows hold H = Gb/V, ghost rows are never read here
    const R *dunit, *linw, *quadw;                // face-indexed copies of the flux metrics [3][sF], [2][sF], [6][sF]
    R* Mb;
    FVM_HD void operator()(int f) const {
        const int sN = m.sN, sF = m.sF, C = m.nInternalCells;
        const int own = m.owner[f], nb = m.neigh[f];
        const bool internal = f < m.nInternalFaces;
        const int kind = face_kind(m, f);
        Geom<R> gm;
        gm.area = m.area[f]; gm.idelta = m.idelta[f];
        for (int k = 0; k < 3; k++) { gm.n[k] = m.normal[(long)k * sF + f]; gm.d[k] = dunit[(long)k * sF + f]; }
        gm.lw[0] = linw[f]; gm.lw[1] = linw[(long)sF + f];
        for (int k = 0; k < 3; k++) { gm.qw[0][k] = quadw[(long)k * sF + f]; gm.qw[1][k] = quadw[(long)(3 + k) * sF + f]; }
        Prim<R> qL, qR; Grad<R> gL, gR;
        load_prim(Q, sN, own, qL); load_grad(G, sN, own, gL); load_prim(Q, sN, nb, qR); load_grad(G, sN, nb, gR);
        // ---- flux: scatter weights A/V, then the flux per unit area
        const R VL = m.vol[own], VR = internal ? m.vol[nb] : R(1);
        R aL[5], aR[5];
        for (int k = 0; k < 5; k++) { aL[k] = abar[(long)k * m.sC + own]; aR[k] = internal ? abar[(long)k * m.sC + nb] : R(0); }
        Flux5<R> F; R wave;
        face_flux(ph, kind, gm, qL, gL, qR, gR, F, wave);
        const R Fk[5] = {F.rho, F.rhoU[0], F.rhoU[1], F.rhoU[2], F.rhoE};
        R sL = R(0), sR = R(0);
        for (int k = 0; k < 5; k++) { sL += Fk[k] * aL[k]; sR += Fk[k] * aR[k]; }
        R Ab = coef * (sL / VL - (internal ? sR / VR : R(0)));
        const R VLb = -coef * gm.area * sL / (VL * VL);
        const R VRb = internal ? coef * gm.area * sR / (VR * VR) : R(0);
        Flux5<R> Fb;
        {
            const R cL = coef * gm.area / VL, cR = internal ? coef * gm.area / VR : R(0);
            Fb.rho = aL[0] * cL - aR[0] * cR; Fb.rhoE = aL[4] * cL - aR[4] * cR;
            for (int k = 0; k < 3; k++) Fb.rhoU[k] = aL[1 + k] * cL - aR[1 + k] * cR;
        }
        Geom<R> gb;
        gb.area = R(0); gb.idelta = R(0);
        for (int k = 0; k < 3; k++) { gb.n[k] = gb.d[k] = gb.qw[0][k] = gb.qw[1][k] = R(0); }
        gb.lw[0] = gb.lw[1] = R(0);
        face_flux_metric_vjp(ph, kind, gm, qL, gL, qR, gR, Fb, gb);
        // ---- Green-Gauss gradient (op.py:45-63): phi_f = w phi_o + (1-w) phi_n, G_o += phi_f S n / V_o, G_n -= phi_f S n / V_n
        R wb = R(0);
        {
            const R w = m.weight[f];
            const R pL[5] = {qL.U[0], qL.U[1], qL.U[2], qL.T, qL.p}, pR[5] = {qR.U[0], qR.U[1], qR.U[2], qR.T, qR.p};
            for (int k = 0; k < 5; k++) {
                const R pf = pL[k] * w + pR[k] * (R(1) - w);
                R hn = R(0);
                for (int j = 0; j < 3; j++) {
                    // row of component k, direction j in Gb: U: 3k+j, T: 9+j, p: 12+j
                    const int row = k < 3 ? 3 * k + j : (k == 3 ? 9 + j : 12 + j);
                    const R dH = Gb[(long)row * sN + own] - (internal ? Gb[(long)row * sN + nb] : R(0));
                    hn += gm.n[j] * dH;
                    gb.n[j] += gm.area * pf * dH;
                }
                Ab += pf * hn;
                wb += gm.area * (pL[k] - pR[k]) * hn;
            }
        }
        // ---- boundary conditions that read the face normal, objective seeds on boundary faces
        if (!internal && f < m.nLocalFaces) {
            const PatchDev<R>& P = m.patches[m.bpatch[f - m.nInternalFaces]];
            Prim<R> q; load_prim(Qb, sN, C + (f - m.nInternalFaces), q);
            if (obja != R(0)) objective_ghost_adj(ph, m, o, Q, obja, f, q);
            if (P.bc[0] == BC_SYMMETRY) {                       // ghost U = u - (u.n) n
                const R un = dot3(qL.U, gm.n), qn = dot3(q.U, gm.n);
                for (int k = 0; k < 3; k++) gb.n[k] += -(qn * qL.U[k] + un * q.U[k]);
            }
            if (P.bc[2] == BC_CBC_TOTAL_PT && !P.dir) {         // direction defaults to the normal (BCs.py:170-176)
                const R Un = dot3(qL.U, gm.n), qd = dot3(q.U, gm.n);
                for (int k = 0; k < 3; k++) gb.n[k] += qL.U[k] * qd + Un * q.U[k] - q.T * Un * qL.U[k] / ph.Cp;
            }
            if (obja != R(0) && (o.kind == OBJ_PATCH_PA || o.kind == OBJ_DRAG) && f >= m.patches[o.patch].startFace &&
                f < m.patches[o.patch].startFace + m.patches[o.patch].nFaces) {
                if (o.kind == OBJ_PATCH_PA) Ab += obja * qR.p;
                else {
                    const R mu = viscosity(ph, qR.T);
                    const R du = qR.U[o.dir] - qL.U[o.dir];
                    Ab += obja * (qR.p * gm.n[o.dir] - mu * du * gm.idelta);
                    gb.n[o.dir] += obja * qR.p * gm.area;
                    gb.idelta += -obja * mu * du * gm.area;
                }
            }
        }
        R* r = Mb + f;
        r[0] += Ab; r[(long)sF] += VLb; r[2L * sF] += VRb; r[3L * sF] += wb;
        r[4L * sF] += -gb.idelta * gm.idelta * gm.idelta;        // deltas = 1 / idelta
        for (int k = 0; k < 3; k++) { r[(long)(5 + k) * sF] += gb.n[k]; r[(long)(8 + k) * sF] += gb.d[k]; }
        r[11L * sF] += gb.lw[0]; r[12L * sF] += gb.lw[1];
        for (int k = 0; k < 3; k++) { r[(long)(13 + k) * sF] += gb.qw[0][k]; r[(long)(16 + k) * sF] += gb.qw[1][k]; }
    }
};
// `volumes` (gradCell divides by it; the cell objective sum T V): Vb_c = - H_c . G_c (+ obja T_c)
template <typename R> struct MeshGradCellBody {
    static constexpr const char* kName = "mesh_grad_cell";
    MeshDev<R> m; const R *Q, *G, *Gb; R objTV; R* Vb;
    FVM_HD void operator()(int c) const {
        R s = R(0);
        for (int k = 0; k < 15; k++) s += Gb[(long)k * m.sN + c] * G[(long)k * m.sN + c];
        Vb[c] += -s + objTV * Q[3L * m.sN + c];
    }
};

// ------------------------------------------------------------------------------------------ layout helpers
// conservative variables of every row (internal + ghost) from the primitives: the tail of the reference's `init` function
// (adFVM/density.py:64-80), used when fields are written with their boundary values
template <typename R> struct ConservativeAllBody {
    static constexpr const char* kName = "conservative_all";
    Phys<R> ph; int sN; const R* Q; R* out;
    FVM_HD void operator()(int i) const {
        Prim<R> q; load_prim(Q, sN, i, q);
        Cons<R> w; conservative(ph, q, w);
        out[i] = w.rho; out[sN + i] = w.rhoU[0]; out[2 * sN + i] = w.rhoU[1]; out[3 * sN + i] = w.rhoU[2]; out[4 * sN + i] = w.rhoE;
    }
};
template <typename R> struct AddBody {             // y += x (objective seeds of the callback objective)
    static constexpr const char* kName = "add";
    R* y; const R* x;
    FVM_HD void operator()(int i) const { y[i] += x[i]; }
};
template <typename R> struct ReciprocalBody {      // x <- 1/x (deltas -> idelta at mesh upload)
    static constexpr const char* kName = "reciprocal";
    R* x;
    FVM_HD void operator()(int i) const { x[i] = R(1) / x[i]; }
};
// AoS host layout ([n][d] row-major, what the reference passes) <-> SoA device layout
// perm (may be NULL = identity): device row i holds host row perm[i] (tile renumbering of cells / internal faces)
template <typename R> struct AosToSoaBody {
    static constexpr const char* kName = "aos_to_soa";
    const R* src; R* dst; int d, stride; const int* perm;
    FVM_HD void operator()(int i) const { const long o = perm ? perm[i] : i; for (int k = 0; k < d; k++) dst[(long)k * stride + i] = src[o * d + k]; }
};
template <typename R> struct SoaToAosBody {
    static constexpr const char* kName = "soa_to_aos";
    const R* src; R* dst; int d, stride; const int* perm;
    FVM_HD void operator()(int i) const { const long o = perm ? perm[i] : i; for (int k = 0; k < d; k++) dst[o * d + k] = src[(long)k * stride + i]; }
};
// processor-patch halo: pack owner rows of the remote faces / unpack into ghost rows (adFVM/cpp/parallel.cpp:116-133)
template <typename R> struct HaloPackBody {
    static constexpr const char* kName = "halo_pack";
    MeshDev<R> m; const R* X; int ncomp; R* buf;
    FVM_HD void operator()(int r) const {
        const int own = m.owner[m.nLocalFaces + r];
        int cs; const long o = halo_offset(m, r, ncomp, cs);
        for (int k = 0; k < ncomp; k++) buf[o + (long)k * cs] = X[k * m.sN + own];
    }
};
template <typename R> struct HaloUnpackBody {
    static constexpr const char* kName = "halo_unpack";
    MeshDev<R> m; R* X; int ncomp; const R* buf;
    FVM_HD void operator()(int r) const {
        int cs; const long o = halo_offset(m, r, ncomp, cs);
        for (int k = 0; k < ncomp; k++) X[k * m.sN + m.nLocalCells + r] = buf[o + (long)k * cs];
    }
};
// reverse halo: ghost-row adjoints of the remote faces -> send buffer
template <typename R> struct HaloPackGhostBody {
    static constexpr const char* kName = "halo_pack_ghost";
    MeshDev<R> m; const R* X; int ncomp; R* buf;
    FVM_HD void operator()(int r) const {
        int cs; const long o = halo_offset(m, r, ncomp, cs);
        for (int k = 0; k < ncomp; k++) buf[o + (long)k * cs] = X[k * m.sN + m.nLocalCells + r];
    }
};

}  // namespace fvm
