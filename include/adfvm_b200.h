/* adfvm_b200 — C ABI of the B200-native adFVM residual + discrete-adjoint hot path.
 *
 * This is the drop-in boundary: what the reference reaches today through its generated CPython module
 * `graph_N.so` (`Function_primal`, `Function_primal_grad`, `initialize`) plus the halo externals of
 * adFVM/cpp/parallel.cpp. Plain pointers and sizes only; all array arguments are HOST pointers in the
 * reference's own layout (C-contiguous row-major AoS `[n][d]`, scalar = float or double chosen at
 * adfvm_create, int32 indices) exactly as `solver.map(*inputs)` passes them (adFVM/solver.py:312-323).
 * Device residency, SoA re-layout, streams and kernels are internal.
 *
 * Every function returns 0 on success; on failure a non-zero code, and adfvm_last_error() describes it
 * (the reference aborts on C asserts / gpuErrorCheck exit, adpy/adpy/cpp/include/gpu.hpp:13-23; raising a
 * Python exception from the binding is strictly better). No call returns partial results.
 * Paths cited below are relative to the reference tree.
 */
#ifndef ADFVM_B200_H
#define ADFVM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct adfvm_ctx adfvm_ctx;

/* enums shared with the Python host layer (adfvm_b200/function.py) */
enum { ADFVM_MU_CONSTANT = 0, ADFVM_MU_SUTHERLAND = 1 };                 /* RCF `mu`, adFVM/density.py:25-31 */
enum { ADFVM_RIEMANN_ROE = 0, ADFVM_RIEMANN_LAXFRIEDRICHS = 1 };          /* adFVM/riemann.py */
enum { ADFVM_PATCH_WALL = 0, ADFVM_PATCH_CYCLIC = 1, ADFVM_PATCH_SYMMETRY = 2, ADFVM_PATCH_EMPTY = 3,
       ADFVM_PATCH_CHARACTERISTIC = 4, ADFVM_PATCH_PROCESSOR = 5, ADFVM_PATCH_PROCESSOR_CYCLIC = 6 };
enum { ADFVM_BC_CALCULATED = 0, ADFVM_BC_CYCLIC = 1, ADFVM_BC_ZEROGRADIENT = 2, ADFVM_BC_FIXEDVALUE = 3,
       ADFVM_BC_SYMMETRY = 4, ADFVM_BC_CBC_UPT = 5, ADFVM_BC_CBC_TOTAL_PT = 6, ADFVM_BC_PROCESSOR = 7 };   /* adFVM/BCs.py */
enum { ADFVM_KEY_VALUE_U = 0, ADFVM_KEY_VALUE_T = 1, ADFVM_KEY_VALUE_P = 2, ADFVM_KEY_U0 = 3, ADFVM_KEY_T0 = 4,
       ADFVM_KEY_P0 = 5, ADFVM_KEY_TT = 6, ADFVM_KEY_PT = 7, ADFVM_KEY_DIRECTION = 8 };                     /* createInput keys */
enum { ADFVM_OBJ_NONE = 0, ADFVM_OBJ_CELL_TV = 1, ADFVM_OBJ_PATCH_PA = 2, ADFVM_OBJ_DRAG = 3,
       ADFVM_OBJ_PLANE_PTLOSS = 4 /* set by adfvm_set_objective_plane */, ADFVM_OBJ_CELL_T = 5 /* sum T, templates/box.py */,
       ADFVM_OBJ_CALLBACK = 6 /* set by adfvm_set_objective_callback */ };
/* option bits of adfvm_primal / adfvm_primal_grad == the kwargs of Function.__call__, adpy/adpy/variable.py:282-287 */
enum { ADFVM_RETURN_STATIC = 1, ADFVM_ZERO_STATIC = 2, ADFVM_REPLACE_STATIC = 4, ADFVM_RETURN_REUSABLE = 8,
       ADFVM_REPLACE_REUSABLE = 16,
       ADFVM_VISCOUS = 32 /* adfvm_primal_grad acts as Function('primal_grad_viscous'), apps/adjoint.py:141,288-289 */ };
/* adjoint artificial viscosity: adjParams[1] of the case file (apps/problem.py:22, adFVM/postpro.py:577-617) */
enum { ADFVM_VISC_NONE = 0, ADFVM_VISC_ABARBANEL = 1, ADFVM_VISC_TURKEL = 2, ADFVM_VISC_UNIFORM = 3 };

/* one boundary patch: the (startFace, nFaces, cellStartFace) triple of Mesh.getScalar (adFVM/mesh.py:874-881) plus
 * what the reference bakes into the generated code from mesh.boundary / field BC dicts at compile time */
typedef struct adfvm_patch {
    int32_t startFace, nFaces, cellStartFace;
    int32_t mesh_type;        /* ADFVM_PATCH_* (mesh.boundary[patch]['type']) */
    int32_t bc_U, bc_T, bc_p; /* ADFVM_BC_*   (CellField.BC classes, adFVM/field.py:123-139) */
    int32_t neighbour_patch;  /* cyclic: index of neighbourPatch in this table, else -1 */
    int32_t peer_rank;        /* processor patches: neighbProcNo, else -1 */
    int32_t tag;              /* processor patches: message ordering tag (adFVM/mesh.py:746-756), else 0 */
} adfvm_patch;

const char* adfvm_last_error(void);
int adfvm_version(void);
/* 1 if this library launches CUDA kernels (the product), 0 for the CPU test simulator built under tests/hostsim */
int adfvm_is_cuda(void);

/* replaces graph_N.initialize(localRank, mesh): adpy/adpy/cpp/module/graph.cpp:16-30 -> cudaSetDevice + external_init.
 * scalar_bytes: 8 (fp64, `--gpu_double`) or 4 (fp32, the reference's GPU default, README.md:87-91).
 * stream: a cudaStream_t to launch on (e.g. torch's current stream) or NULL for the default stream. */
int adfvm_create(adfvm_ctx** ctx, int device, int scalar_bytes, void* stream);
int adfvm_destroy(adfvm_ctx* ctx);

/* RCF constants (adFVM/density.py:22-39) */
int adfvm_set_physics(adfvm_ctx* ctx, double gamma, double Cp, double Pr, int mu_law, double mu_value,
                      int riemann_solver, int boundary_riemann_solver);

/* static inputs 4..26(+3/patch) of `primal`: Mesh.gradFields, Mesh.intFields, Mesh.constants, patch table.
 * sizes = {nCells, nFaces, nInternalCells, nInternalFaces, nLocalCells, nRemoteCells, nLocalFaces, nGhostCells}.
 * Uploaded once (the reference's "static" arrays, adpy/adpy/cpp/include/common.hpp:315-327). */
int adfvm_set_mesh(adfvm_ctx* ctx, const int32_t sizes[8],
                   const void* areas, const void* volumesL, const void* volumesR, const void* weights,
                   const void* deltas, const void* normals, const void* deltasUnit, const void* linearWeights,
                   const void* quadraticWeights, const void* volumes,
                   const int32_t* owner, const int32_t* neighbour, const int32_t* cellFaces,
                   const int32_t* cellNeighbours, const int32_t* cellOwner,
                   int32_t nPatches, const adfvm_patch* patches);
/* BC value arrays ([nFaces][d], static inputs after the source terms; adFVM/BCs.py:45-54). May be called again
 * later: this is the explicit "invalidate static data" the reference lacks on GPU (apps/problem.py:109-110). */
int adfvm_set_bc_value(adfvm_ctx* ctx, int32_t patch, int32_t key, const void* values);
/* objective evaluated on the stage-1 state (adFVM/density.py:412-413); see DESIGN.md for the supported kinds */
int adfvm_set_objective(adfvm_ctx* ctx, int32_t kind, int32_t patch, int32_t direction);
/* parameter block of the adjoint (reference apps/adjoint.py:101-120): the source terms (default, patch < 0) or ONE
 * boundary-condition input array - parameters = ('BCs', field, patch, key) - whose gradient [nFaces][d] adfvm_primal_grad
 * then returns in its first gradient array (same static-accumulator options). key: ADFVM_KEY_* of adfvm_set_bc_value. */
int adfvm_set_parameter_bc(adfvm_ctx* ctx, int32_t patch, int32_t key);
/* parameters = 'mesh' (apps/adjoint.py:105-107): gradient with respect to the ten metric arrays of adFVM/mesh.py:27-31.
 * Select it BEFORE adfvm_set_mesh; after adjoint calls read the accumulated gradients (reference layout and numbering:
 * areas [F], volumesL [F], volumesR [Fi], weights [F], deltas [F], normals [F][3], deltasUnit [F][3], linearWeights
 * [F][2], quadraticWeights [F][2][3], volumes [C]) with adfvm_get_mesh_grad. */
int adfvm_set_parameter_mesh(adfvm_ctx* ctx);
int adfvm_get_mesh_grad(adfvm_ctx* ctx, void* areas, void* volumesL, void* volumesR, void* weights, void* deltas, void* normals,
                        void* deltasUnit, void* linearWeights, void* quadraticWeights, void* volumes, int32_t zero_static);
/* the design objective of the reference's turbine-vane cases (adFVM/objectives/vane.py:36-66,83-139, templates/vane.py):
 * scale x mass-flow averaged total-pressure loss (ptin - pt)/ptin over the cells of a cut plane. cells (reference
 * numbering) and areas are the extraArgs the case file passes after the BC arrays (adFVM/solver.py:317). */
int adfvm_set_objective_plane(adfvm_ctx* ctx, int32_t n, const int32_t* cells, const void* areas, double ptin,
                              const double normal[3], double scale);
/* == Function_init (adFVM/density.py:64-80; called by RCF.initFields / writeFields): conservative fields WITH ghost rows -
 * primitive, ghost fill by the boundary conditions and the halo, conservative on every row. In: state [C][d]; out: host arrays
 * [nCells][d] in the reference's numbering (internal cells, then one ghost row per boundary face). The resident state of
 * adfvm_primal is not touched. */
int adfvm_init_fields(adfvm_ctx* ctx, const void* rho, const void* rhoU, const void* rhoE,
                      void* rho_out, void* rhoU_out, void* rhoE_out);
/* ANY objective of a case file (the reference differentiates arbitrary adpy-DSL kernels, templates/cylinder_test.py:9-36,
 * adFVM/objectives/vane.py:83-140, traced by adpy/adpy/tensor.py:444-485): the host layer evaluates the traced kernels and their
 * reverse mode on the device arrays (adfvm_b200/adpy_objective.py) through this callback. It is called on the stage-1 primitives
 * Q = [5][stride] (U_x,U_y,U_z,T,p; rows [0,nInternalCells) in DEVICE cell order - see adfvm_get_cell_perm - ghost rows in the
 * reference's order) and returns the rank-local objective; with want_seed it also writes obja * dObjective/dQ into Qseed
 * (same layout, zeroed by the library). Pointers are device pointers, work must be ordered on the context's stream. */
typedef double (*adfvm_objective_fn)(void* user, const void* Q, int64_t stride, int32_t want_seed, double obja, void* Qseed);
int adfvm_set_objective_callback(adfvm_ctx* ctx, adfvm_objective_fn fn, void* user);
/* device cell order: perm[i] = reference (host) index of the cell in device row i, i < nInternalCells */
int adfvm_get_cell_perm(adfvm_ctx* ctx, int32_t* perm);
/* source terms [C][1],[C][3],[C][1] (Solver.sourceTerms, adFVM/solver.py:116-133); static, re-settable */
int adfvm_set_source(adfvm_ctx* ctx, const void* S_rho, const void* S_rhoU, const void* S_rhoE);

/* == Function_primal (adpy/adpy/variable.py:360-504 as instantiated by adFVM/density.py:101-105) ==
 * rho,rhoU,rhoE: state in; read only if ADFVM_REPLACE_REUSABLE or no state is resident yet (reuse ids primal_0..2).
 * *_out: written only if ADFVM_RETURN_REUSABLE (may be NULL otherwise). dtc/obj: one scalar each, always written
 * (dtc = rank-local max over cells, obj = objective summed over ranks), in the context's scalar type. */
int adfvm_primal(adfvm_ctx* ctx, const void* rho, const void* rhoU, const void* rhoE, double dt, int32_t options,
                 void* rho_out, void* rhoU_out, void* rhoE_out, void* dtc_out, void* obj_out);

/* == Function_primal_grad for parameters='source' (apps/adjoint.py:94-126, 268-291) ==
 * state at the START of the step, adjoint of the step OUTPUT (volume-weighted, as the reference carries it),
 * dtca (ignored: max-reduce has no gradient, adpy/adpy/scalar.py:306-311), obja.
 * Always writes the new adjoint fields; the source-term gradients are static accumulators: written to grad_*
 * only with ADFVM_RETURN_STATIC and zeroed only with ADFVM_ZERO_STATIC (adpy/adpy/variable.py:484-490). */
int adfvm_primal_grad(adfvm_ctx* ctx, const void* rho, const void* rhoU, const void* rhoE, double dt,
                      const void* rhoa, const void* rhoUa, const void* rhoEa, double dtca, double obja, int32_t options,
                      void* rhoa_out, void* rhoUa_out, void* rhoEa_out,
                      void* grad_S_rho, void* grad_S_rhoU, void* grad_S_rhoE);

/* device-resident stepping (no host transfers): what Solver.run does between report steps (return_reusable=0) */
int adfvm_primal_step_resident(adfvm_ctx* ctx, double dt);
/* == adjoint artificial viscosity: Function('primal_grad_viscous') (apps/adjoint.py:127-141) ==
 * Replaces computeAdjointViscosity (adFVM/postpro.py:491-697), Function_get_max_eigenvalue (adFVM/cpp/scaling.cpp:23-62 cusolver
 * syevjBatched / :84-106 LAPACK dsyev), Function_apply_adjoint_viscosity (scaling.cpp:134-160) and Matop::heat_equation
 * (adFVM/cpp/matop_cuda.cpp:180-235, matop_petsc.cpp:230-484). type: ADFVM_VISC_*; scaling: adjParams[0] (the `scaling` input of
 * primal_grad); rtol / maxit of the diffusion solve (<= 0: 1e-13 fp64 / 1e-6 fp32, 500). Afterwards adfvm_primal_grad with
 * ADFVM_VISCOUS smooths the adjoint fields it returns: M_2norm from the step's start state, one implicit diffusion step over dt. */
int adfvm_set_adjoint_viscosity(adfvm_ctx* ctx, int32_t type, double scaling, double rtol, int32_t maxit);
/* the same smoothing applied to the resident adjoint fields (after adfvm_adjoint_step_resident / adfvm_adjoint_block) */
int adfvm_adjoint_viscous_resident(adfvm_ctx* ctx, double dt);
/* adfvm_adjoint_block with the smoothing applied after every step (viscousInterval = 1, apps/adjoint.py:250,288-289) */
int adfvm_adjoint_block_viscous(adfvm_ctx* ctx, int32_t nsteps, const double* dt, double obja);
/* diagnostic (write_M_2norm, apps/adjoint.py:26,139): M_2norm [nCells][1] of a host state, ghost rows filled; reference numbering */
int adfvm_get_adjoint_viscosity(adfvm_ctx* ctx, const void* rho, const void* rhoU, const void* rhoE, void* M_2norm);
/* conjugate-gradient iterations of the last diffusion solve */
int64_t adfvm_viscosity_iterations(adfvm_ctx* ctx);
/* == opt-in: device copies of states handed out to the host (new; the reference uploads the state at every primal_grad call,
 * apps/adjoint.py:268-280) == reserve n slots; adfvm_state_cache_put(key) files the resident state under `key` (least recently
 * used slot replaced); adfvm_state_cache_select(key, &found) makes the NEXT adfvm_primal (with ADFVM_REPLACE_REUSABLE) or
 * adfvm_primal_grad call that passes rho = rhoU = rhoE = NULL start from that copy. The host layer (PrimalFunction(state_cache=n))
 * selects a key only for the very array objects it returned, which it hands out read-only. */
int adfvm_state_cache_reserve(adfvm_ctx* ctx, int32_t nslots);
int adfvm_state_cache_put(adfvm_ctx* ctx, int64_t key);
int adfvm_state_cache_select(adfvm_ctx* ctx, int64_t key, int32_t* found);
int64_t adfvm_state_cache_hits(adfvm_ctx* ctx);
/* adjoint step on resident data; chain!=0 feeds the previous call's output adjoint back in as this call's input */
int adfvm_adjoint_step_resident(adfvm_ctx* ctx, double dt, double obja, int32_t chain);
int adfvm_get_dtc_obj(adfvm_ctx* ctx, double* dtc, double* obj);
/* the same dtc maximised over all ranks of the communicator (a collective: every rank calls it once per step). This is the
 * `parallel.min(2*CFL/dtc)` of the adaptive time step, adFVM/solver.py:365, for launches without mpi4py (torchrun + NCCL):
 * dt_next = min(2*CFL/dtc_global, dt*stepFactor, endTime - t). */
int adfvm_get_dtc_global(adfvm_ctx* ctx, double* dtc);
int adfvm_get_state(adfvm_ctx* ctx, void* rho, void* rhoU, void* rhoE);
int adfvm_sync(adfvm_ctx* ctx);
/* kernels launched by this context so far */
int64_t adfvm_launch_count(adfvm_ctx* ctx);
/* bytes of device memory held by the context */
int64_t adfvm_device_bytes(adfvm_ctx* ctx);

/* Tiling of the flux kernels (internal data layout): cells are regrouped into tiles of `cells`
 * consecutive cells, one CTA per tile. adfvm_set_tile_cells (64 or 128; default 128, with automatic fallback to 64 on 1-D/2-D meshes) must precede adfvm_set_mesh.
 * adfvm_tile_stats reports flux evaluations per cell (3 = every face once, 6 = cell-centred gather), the largest
 * number of rounds of a 32-cell sub-tile (4 on a regular hex block), the number of tiles and the tile size. */
int adfvm_set_tile_cells(adfvm_ctx* ctx, int32_t cells);
int adfvm_tile_stats(adfvm_ctx* ctx, double* evals_per_cell, int32_t* max_rounds, int32_t* n_tiles, int32_t* tile_cells);

/* largest halo (slots beyond the tile's own cells) of any tile, the kernel variant chosen from it and a
 * histogram of tiles per halo size in bins of 32 slots */
int adfvm_tile_halo_stats(adfvm_ctx* ctx, int32_t* max_halo, int32_t* variant, int32_t* hist, int32_t nbins);

/* Mesh metric build on the device: replaces cmesh.build (adFVM/cpp/cmesh.cpp:65-270, called from adFVM/mesh.py:107) and the
 * ghost-cell centres of Mesh.createGhostCells (adFVM/mesh.py:758-819). Host arrays in the reference's AoS layout:
 * points [nPoints][3], faces [nFaces][4] (quads), owner [nFaces], neighbour [nFaces] (boundary face b -> ghost cell
 * nInternalCells + b), cellFaces [nInternalCells][6]. patches: kind 0 plain (ghost centre = face centre), 1 cyclic
 * (nbrStartFace = first face of neighbourPatch), 2 processor (centres taken from remote_centres [nBoundaryFaces][3]).
 * Outputs: areas [nF], normals [nF][3], faceCentres [nF][3], cellCentres [nCells][3], volumes [nIC], deltas [nF],
 * deltasUnit [nF][3], weights [nF], linearWeights [nF][2], quadraticWeights [nF][2][3]. */
typedef struct { int32_t startFace, nFaces, kind, nbrStartFace; } adfvm_metric_patch;
int adfvm_mesh_metrics(adfvm_ctx* ctx, int32_t nPoints, const void* points, int32_t nFaces, int32_t nInternalFaces,
                       int32_t nInternalCells, const int32_t* faces, const int32_t* owner, const int32_t* neighbour,
                       const int32_t* cellFaces, int32_t nPatches, const adfvm_metric_patch* patches, const void* remote_centres,
                       void* areas, void* normals, void* faceCentres, void* cellCentres, void* volumes, void* deltas,
                       void* deltasUnit, void* weights, void* linearWeights, void* quadraticWeights);

/* Device-side checkpoint block: what Solver.run(mode='forward') + the step loop of Adjoint.run do through host memory
 * (adFVM/solver.py:376-382: every state of a block returned to numpy; apps/adjoint.py:217-291: fed back one by one).
 * adfvm_primal_block runs nsteps steps from the resident state with time steps dt[k], keeps the state at the start of
 * every step in HBM and returns the per-step dtc / objective. adfvm_adjoint_block then walks that block backwards from
 * the resident adjoint fields (set by adfvm_set_adjoint, or left by a previous adjoint call/block; objective seed obja
 * per step); adfvm_get_adjoint reads the adjoint fields (and, if g_* are given, the accumulated source-term gradient). */
int adfvm_set_state(adfvm_ctx* ctx, const void* rho, const void* rhoU, const void* rhoE);   /* upload the resident state without stepping */
int adfvm_primal_block(adfvm_ctx* ctx, int32_t nsteps, const double* dt, double* dtc, double* objective);
int adfvm_adjoint_block(adfvm_ctx* ctx, int32_t nsteps, const double* dt, double obja);
int adfvm_set_adjoint(adfvm_ctx* ctx, const void* rhoa, const void* rhoUa, const void* rhoEa);
int adfvm_get_adjoint(adfvm_ctx* ctx, void* rhoa, void* rhoUa, void* rhoEa, void* g_rho, void* g_rhoU, void* g_rhoE, int32_t zero_static);

/* whole steps are captured into CUDA graphs (the second time a step with the same time step and buffer rotation runs)
 * and replayed afterwards; number of steps served by a replay so far (set ADFVM_NO_GRAPH=1 to run every step eagerly) */
int64_t adfvm_graph_replays(adfvm_ctx* ctx);

/* rounds (32 lanes, one face evaluation each) summed over all sub-tiles, the number of sub-tiles, and the number of
 * tiles that do not depend on processor-patch data (the range that overlaps the halo exchange) */
int adfvm_tile_rounds(adfvm_ctx* ctx, int64_t* rounds, int64_t* subtiles, int32_t* early_tiles);

/* per-kernel device timing (CUDA events on the launching stream around every launch while enabled).
 * adfvm_kernel_report writes lines "<kernel> <launches> <total_ms>\n" into buf. Counterpart of the reference's
 * `-o/--profile` per-kernel prints (adpy/adpy/variable.py:437-467). */
int adfvm_kernel_timing(adfvm_ctx* ctx, int32_t enable);
int adfvm_kernel_report(adfvm_ctx* ctx, char* buf, int32_t buflen);

/* page-locked host memory for the arrays a call returns: replaces putArray's per-call `new T[size]`
 * (adpy/adpy/cpp/include/interface.hpp:54-80); device-to-host copies into it are asynchronous and run at the full
 * PCIe rate. The Python host layer recycles the blocks when the numpy arrays that wrap them are released. */
int adfvm_host_alloc(void** out, size_t bytes);
int adfvm_host_free(void* p);

/* multi-GPU halo: replaces Function_mpi_init / Function_mpi / Function_mpi_end and their _grad twins
 * (adFVM/cpp/parallel.cpp:31-209) and Function_mpi_allreduce (:214-232) with NCCL point-to-point over NVLink.
 * id: 128-byte ncclUniqueId produced on rank 0 and distributed by the host layer (torch.distributed). */
int adfvm_comm_unique_id(void* id128);
int adfvm_comm_init(adfvm_ctx* ctx, const void* id128, int32_t rank, int32_t nranks);

#ifdef __cplusplus
}
#endif
#endif
