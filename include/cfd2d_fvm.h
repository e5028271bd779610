/* cfd2d_fvm.h -- C-ABI of the B200 (sm_100a) explicit finite-volume path of cfd-2d.
 *
 * Drop-in boundary.  The reference (zhrv/cfd-2d) has no FFI: its solver "plugin" interface is the
 * abstract C++ class Method { init(char* xml); run(); done(); } (src/methods/method.h:6-135),
 * implemented for this path by class FVM_TVD (src/methods/fvm_tvd.{h,cpp}).  The entry points
 * below are exactly what a Method subclass needs to hand the hot loop to the GPU; each one cites
 * the reference code it replaces.  The glue subclass (cfd-2d_b200/host/fvm_tvd_cuda.cpp) and the
 * one-line factory hook are shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the handle owns all
 *     device memory; one handle per GPU (per rank); a handle is not thread-safe, independent
 *     handles may be driven from different threads;
 *   - every function returns 0 on success or a negative CFD2D_E* code, never exit()s
 *     (the reference log()+exit()s, e.g. fvm_tvd.cpp:16-17,150-151); the message is available
 *     from cfd2d_fvm_last_error();
 *   - all arithmetic is FP64, indices are int32 (the reference's int), cell flags are uint32
 *     (Cell::flag, src/mesh/grid.h:36);
 *   - there is NO CPU fallback: create() fails with CFD2D_ENODEV when no CUDA device is usable.
 */
#ifndef CFD2D_FVM_H
#define CFD2D_FVM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFD2D_OK        0
#define CFD2D_EINVAL   (-1)   /* bad argument / inconsistent mesh                                  */
#define CFD2D_ENODEV   (-2)   /* no usable CUDA device (there is no CPU fallback)                  */
#define CFD2D_ECUDA    (-3)   /* CUDA runtime error (see cfd2d_fvm_last_error)                     */
#define CFD2D_ENEWTON  (-4)   /* rim_orig Newton iteration hit the cap (the reference would hang,  */
                              /* src/global.cpp:278-307 has no cap)                                */
#define CFD2D_EBC      (-5)   /* boundary edge without boundary condition (fvm_tvd.cpp:706-710)    */
#define CFD2D_ENCCL    (-6)   /* NCCL error / NCCL not loadable for a multi-rank handle            */

/* boundary kinds: CFDBndInlet / CFDBndOutlet / CFDBndWallSlip (src/bnd_cond.cpp:75-110);
 * BOUND_WALL_NO_SLIP instantiates the slip class (bnd_cond.cpp:50-55) => same kind.              */
#define CFD2D_BC_INLET   1
#define CFD2D_BC_OUTLET  2
#define CFD2D_BC_WALL    3

/* numerical flux: the live Godunov/exact-Riemann block of FVM_TVD::calcFlux (fvm_tvd.cpp:604-622,
 * rim_orig global.cpp:232-405) or the commented Lax-Friedrichs block (fvm_tvd.cpp:623-642).      */
#define CFD2D_FLUX_GODUNOV 0
#define CFD2D_FLUX_LAX     1

#define CFD2D_CELL_FLAG_LIM 0x000002u   /* src/mesh/grid.h:14 */

/* Flattened Grid (src/mesh/grid.h:18-100) of ONE rank.  Cells [0,nc) are owned, [nc,nc_ex) are
 * halo cells grouped by owner rank ascending (Grid::cCount / cCountEx, grid.cpp:423-430).  Every
 * edge keeps the GLOBAL orientation (c1 -> c2, normal out of c1) of the undecomposed mesh.       */
typedef struct cfd2d_mesh {
    int32_t nc;                 /* owned cells                     Grid::cCount                   */
    int32_t nc_ex;              /* owned + halo cells              Grid::cCountEx (== nc serial)   */
    int32_t ne;                 /* edges touching an owned cell                                    */
    const double*  cell_S;      /* [nc_ex] Cell::S                                                 */
    const double*  cell_cx;     /* [nc_ex] Cell::c.x                                               */
    const double*  cell_cy;     /* [nc_ex] Cell::c.y                                               */
    const int32_t* cell_mat;    /* [nc_ex] material index of the cell's region (fvm_tvd.cpp:788)   */
    const int32_t* cell_edges;  /* [3*nc]  Cell::edgesInd, ascending edge id (the summation order) */
    const int32_t* edge_c1;     /* [ne]    Edge::c1                                                */
    const int32_t* edge_c2;     /* [ne]    Edge::c2, -1 on a boundary edge                         */
    const double*  edge_nx;     /* [ne]    Edge::n.x                                               */
    const double*  edge_ny;     /* [ne]    Edge::n.y                                               */
    const double*  edge_l;      /* [ne]    Edge::l                                                 */
    const double*  edge_gp;     /* [4*ne]  Edge::c[1].x, c[1].y, c[2].x, c[2].y (2 Gauss points)   */
    const int32_t* edge_bc;     /* [ne]    index into the BC table, -1 on inner edges (Edge::bnd)  */
} cfd2d_mesh;

/* Materials (global.h:199-222), boundary table (bnd_cond.h:14-35), limits (fvm_tvd.cpp:43-48).   */
typedef struct cfd2d_phys {
    int32_t nmat;
    const double*  mat_M;       /* [nmat] Material::M                                              */
    const double*  mat_Cp;      /* [nmat] Material::Cp                                             */
    int32_t nbc;
    const int32_t* bc_kind;     /* [nbc]  CFD2D_BC_*                                               */
    const double*  bc_par;      /* [4*nbc] inlet: Vx, Vy, T, P (CFDBoundary::par)                  */
    double limits[5];           /* limitRmin, limitRmax, limitPmin, limitPmax, limitUmax           */
} cfd2d_phys;

/* <control> of task.xml (fvm_tvd.cpp:26-40) + the scheme selectors.                              */
typedef struct cfd2d_ctrl {
    double  CFL;
    double  TAU;
    int32_t steady;             /* STEADY: 1 = local time step recomputed every step               */
    int32_t flux;               /* CFD2D_FLUX_*            (reference live code: GODUNOV)          */
    int32_t order;              /* 2 = Green-Gauss linear reconstruction (reference live code),    */
                                /* 1 = piecewise constant (the "//return;" of reconstruct :654)    */
    int32_t max_newton;         /* cap on rim_orig's Newton loop, <=0 -> 1000                      */
} cfd2d_ctrl;

/* Halo description of one rank (Grid::recvCount / recvShift / sendInd, grid.h:94-96,
 * filled by Decomp, src/methods/decomp.cpp:172-292).  NULL for a serial handle.                  */
typedef struct cfd2d_halo {
    int32_t rank, nranks;
    const int32_t* recv_count;  /* [nranks] halo cells owned by rank p; they sit contiguously at   */
                                /*          nc + sum(recv_count[0..p))  (Grid::recvShift)          */
    const int32_t* send_count;  /* [nranks] sendInd[p].size()                                      */
    const int32_t* send_ind;    /* [sum send_count] owned-cell indices, rank-major (Grid::sendInd) */
    const void*    nccl_unique_id; /* 128-byte ncclUniqueId, identical on every rank               */
    const int32_t* cell_gid;    /* [nc_ex] optional (may be NULL): global id of every local cell   */
                                /* (Decomp's cell map, decomp.cpp:165-211).  Used to verify that   */
                                /* every edge across the partition keeps id(c1) < id(c2), which    */
                                /* makes the per-rank remediateLimCells equal the serial ascending  */
                                /* sweep (fvm_tvd.cpp:464-499)                                      */
} cfd2d_halo;

typedef struct cfd2d_fvm cfd2d_fvm;

/* Replaces the allocation part of FVM_TVD::init (fvm_tvd.cpp:177-197): uploads the flattened mesh,
 * BC/material tables and control block to `device`, derives the per-cell gather tables.          */
int cfd2d_fvm_create(const cfd2d_mesh* mesh, const cfd2d_phys* phys, const cfd2d_ctrl* ctrl,
                     const cfd2d_halo* halo, int device, cfd2d_fvm** out);

/* Replaces FVM_TVD::done (fvm_tvd.cpp:731-752). NULL is allowed.                                  */
void cfd2d_fvm_destroy(cfd2d_fvm* h);

/* Replaces the initial-state loop + copy to *_old (fvm_tvd.cpp:199-209): conservative variables of
 * the OWNED cells [nc]; flag may be NULL (= all zero; the reference leaves Cell::flag
 * uninitialised).                                                                  */
int cfd2d_fvm_set_state(cfd2d_fvm* h, const double* ro, const double* ru, const double* rv,
                        const double* re, const uint32_t* flag);

/* Replaces FVM_TVD::calcTimeStep (fvm_tvd.cpp:216-240): unsteady -> TAU = min(TAU, min_c CFL*S/
 * max(|u|+c,|v|+c)) (all-reduced over ranks), cTau[] = TAU; steady -> per-cell cTau.             */
int cfd2d_fvm_calc_time_step(cfd2d_fvm* h, double* tau_out);

/* Replaces the body of the while loop of FVM_TVD::run (fvm_tvd.cpp:310-450) for nsteps whole RK2
 * steps: copy to old, 2 x (gradients, edge fluxes, residual gather, update), half-sum, limit
 * flags, remediateLimCells.  Returns after the work has completed on the device.                 */
int cfd2d_fvm_step(cfd2d_fvm* h, int nsteps);

/* Same, but only enqueues on the handle's stream (pair with cfd2d_fvm_sync).                      */
int cfd2d_fvm_step_async(cfd2d_fvm* h, int nsteps);
int cfd2d_fvm_sync(cfd2d_fvm* h);

/* What FVM_TVD::save / the log line need (fvm_tvd.cpp:452-459, :501-600): conservative state of
 * the owned cells; cTau and flag may be NULL.                                                     */
int cfd2d_fvm_get_state(cfd2d_fvm* h, double* ro, double* ru, double* rv, double* re,
                        double* cTau, uint32_t* flag);

/* The same without stalling the time loop (FVM_TVD::run saves every FILE_OUTPUT_STEP steps,
 * fvm_tvd.cpp:452-455, and its ASCII VTK writer :501-600 dominates wall time at large N):
 * snapshot_begin captures the state as of the steps enqueued so far and starts the device-to-host
 * copy on a separate copy stream; the caller enqueues the next chunk (cfd2d_fvm_step_async), then
 * snapshot_end waits for the copy only and fills the arrays (any pointer may be NULL) -- the file is
 * written while the GPU runs the next chunk.  One snapshot may be outstanding.                      */
int cfd2d_fvm_snapshot_begin(cfd2d_fvm* h);
int cfd2d_fvm_snapshot_end(cfd2d_fvm* h, double* ro, double* ru, double* rv, double* re,
                           double* cTau, uint32_t* flag);

/* Multi-rank handles: collect the owned-cell state of every rank on rank `root` over NCCL -- what a
 * parallel Method does with Parallel::send/recv (global.cpp:607-659) before its root rank writes the
 * result file.  counts[nranks] = owned cells of each rank (collective: every rank calls it with the
 * same root and counts); on root the arrays (size sum(counts), any may be NULL) receive the ranks'
 * blocks one after the other, each in its rank's own cell order; other ranks may pass NULL arrays.   */
int cfd2d_fvm_gather_state(cfd2d_fvm* h, int root, const int32_t* counts, double* ro, double* ru,
                           double* rv, double* re, double* cTau, uint32_t* flag);

/* Primitive fields FVM_TVD::save prints (convertConsToPar per cell, fvm_tvd.cpp:529-572),
 * converted on the device; any pointer may be NULL.                                               */
int cfd2d_fvm_get_primitive(cfd2d_fvm* h, double* r, double* p, double* T, double* u, double* v,
                            double* cz);

/* Current TAU (unsteady) and simulated time t (fvm_tvd.cpp:313).                                  */
double cfd2d_fvm_tau(const cfd2d_fvm* h);
double cfd2d_fvm_time(const cfd2d_fvm* h);

/* Parity hooks: one evaluation of FVM_TVD::calcGrad (fvm_tvd.cpp:242-301) -> grad8[nc][8] =
 * (Rx,Ry,Px,Py,Ux,Uy,Vx,Vy); and of the edge-flux sweep (fvm_tvd.cpp:329-352) -> flux4[ne][4] =
 * (fr,fu,fv,fe) summed over the two Gauss points (before the l/2 factor).                         */
int cfd2d_fvm_calc_grad(cfd2d_fvm* h, double* grad8);
int cfd2d_fvm_edge_fluxes(cfd2d_fvm* h, double* flux4);

/* Function-level known-answer entry points (run on `device`):
 *   rim_orig (global.cpp:232-405): in8[n][8] = RB,PB,UB,VB,RE,PE,UE,VE -> out5[n][5] =
 *   RI,EI,PI,UI,VI; iters[n] (may be NULL) = Newton iterations taken.                             */
int cfd2d_kat_rim_orig(int device, int n, const double* in8, double gam, int max_newton,
                       double* out5, int32_t* iters);
/*   the same through the reduced-instruction solver the Godunov kernels use by default
 *   (cfd2d_fvm_use_exact_riemann); gamma is the flux loop's hard-coded 1.4 (fvm_tvd.cpp:345).     */
int cfd2d_kat_rim_orig_fast(int device, int n, const double* in8, int max_newton, double* out5,
                            int32_t* iters);
/*   FVM_TVD::calcFlux (fvm_tvd.cpp:602-643): in12[n][12] = rL,pL,uL,vL,EL, rR,pR,uR,vR,ER, nx,ny;
 *   flux = CFD2D_FLUX_GODUNOV (bit-faithful rim_orig), CFD2D_FLUX_LAX, or 2 = Godunov through the
 *   reduced-instruction solver.                                                                    */
int cfd2d_kat_calc_flux(int device, int n, const double* in12, double gam, int flux, double* out4);

/*   Material::URS (global.cpp:9-30): io8[n][8] = r,p,e,E,u,v,cz,T updated in place; mode 0: (r,e) ->
 *   p,cz; mode 1: (r,p) -> e,T; mode 2: (p,T) -> r,cz -- the three uses on this path
 *   (convertConsToPar fvm_tvd.cpp:803-813, boundaryCond :694-711).                                 */
int cfd2d_kat_urs(int device, int n, double M, double Cp, int mode, double* io8);

/* Per-kernel device time: runs nsteps steps with CUDA events around every launch.
 * ms[CFD2D_NKERNELS] receives the summed milliseconds per kernel, launches[] the launch counts.   */
#define CFD2D_K_GRAD      0   /* K2: BC ghost + Green-Gauss gradient gather                         */
#define CFD2D_K_FLUX      1   /* K3: reconstruction + numerical flux per edge                       */
#define CFD2D_K_UPDATE1   2   /* K4: residual gather + RK stage-1 update                            */
#define CFD2D_K_UPDATE2   3   /* K5: residual gather + stage-2 update + half-sum + limit flags      */
#define CFD2D_K_REMEDIATE 4   /* K6: remediateLimCells                                              */
#define CFD2D_K_TIMESTEP  5   /* K1: steady local time step                                         */
#define CFD2D_K_HALO      6   /* K7: halo pack + exchange                                           */
#define CFD2D_K_STAGE1    7   /* whole RK stage 1 in one kernel: tile-fused k_stage (use_fused) or the  */
                              /* single-sweep first-order Lax-Friedrichs kernel k_cell_lf1            */
#define CFD2D_K_STAGE2    8   /* same for stage 2 (+ half-sum + limit flags)                          */
#define CFD2D_NKERNELS    9
int cfd2d_fvm_profile(cfd2d_fvm* h, int nsteps, double* ms, int64_t* launches);

/* Kernel launches issued by this handle since create (the bench's gpu_launches claim).            */
int64_t cfd2d_fvm_launch_count(const cfd2d_fvm* h);

/* How this handle's Method::exchange (method.h:13-41) moves halo records: 0 serial handle, 1 NCCL
 * send/recv, 2 direct peer stores over NVLink (CUDA IPC; CFD2D_HALO_P2P=1 at create on every rank,
 * falls back to 1 when a neighbour cannot be mapped).                                             */
int cfd2d_fvm_halo_transport(const cfd2d_fvm* h);

/* Use an externally created cudaStream_t (e.g. the caller's current stream) for all launches.     */
int cfd2d_fvm_set_stream(cfd2d_fvm* h, void* cuda_stream);

/* Enable/disable CUDA-graph replay of the step (default on).                                      */
int cfd2d_fvm_use_graph(cfd2d_fvm* h, int on);

/* Choose how the step is laid out on the device: 0 = three sweeps per stage as in FVM_TVD::run
 * (gradient, edge-flux and update kernels with HBM staging); 1 = one tile-fused kernel per RK stage
 * (k_stage: gradients and edge fluxes stay in shared memory); 2 = the persistent, software-pipelined
 * tile kernel (k_stage_pipe: per-tile tables and state ranges arrive by cp.async.bulk + mbarrier one
 * tile ahead, the primitive cache leaves HBM).  All three produce the same bits.  CFD2D_FUSED=0|1|2
 * selects the layout at create().                                                                  */
int cfd2d_fvm_use_fused(cfd2d_fvm* h, int on);

/* Godunov handles only.  0 (default): rim_orig (global.cpp:232-405) is evaluated by the reduced-
 * instruction solver (shared reciprocals, FMA, x^(1/7) by Newton; same algorithm, same branches,
 * results within a few ulp per operation -- inside the 1e-12 contract of this exp/log path);
 * 1: by the statement that keeps the reference's operation order (CFD2D_EXACT_RIEMANN=1 at create).  */
int cfd2d_fvm_use_exact_riemann(cfd2d_fvm* h, int on);

/* One-line description of the tile plan of this handle (tiles, ring overhead, shared memory).     */
const char* cfd2d_fvm_plan_summary(const cfd2d_fvm* h);

/* Host-only test hook (no GPU needed): builds the cell renumbering and the tile plan create()
 * would build for `mesh` and verifies its invariants (every cell slot finds its edge in the tile,
 * local ids consistent, ring-1 closure).  perm_out[nc_ex] (may be NULL) = caller -> device cell
 * id; stats_out[8] (may be NULL) = ntiles, nl_max, ne_max, sum n_g, sum ne_t, sum ring-1,
 * interior tiles, boundary tiles.  Returns 0 or CFD2D_EINVAL (message via last_error(NULL)).      */
int cfd2d_tiling_plan(const cfd2d_mesh* mesh, int tile_cells, int hilbert, int32_t* perm_out,
                      int64_t* stats_out);

/* Host-only test hook for the plan of the pipelined tile kernel (layout 2): builds the per-tile blobs
 * create() would build and re-derives every table from the blob bytes (slots -> edges -> local cell
 * ids, geometry, ring-1 gradient tables, materials).  stats_out[12] (may be NULL) = ntiles, max
 * staged cells, max edges, max blob bytes, sum edges, sum ring 1, sum ring 2, blob bytes, interior
 * tiles, boundary tiles, max gathered records, max cells with a gradient.                          */
int cfd2d_pipe_plan(const cfd2d_mesh* mesh, int tile_cells, int dir_bins, int hilbert, int64_t* stats_out);

/* Multi-rank bootstrap: a 128-byte ncclUniqueId created on one rank (ncclGetUniqueId); the caller
 * ships it to the other ranks (MPI_Bcast in the reference host, torch.distributed here) and every
 * rank passes it in cfd2d_halo.nccl_unique_id.                                                    */
int cfd2d_nccl_get_unique_id(void* out128);

/* ---- mesh ingest (host only) -------------------------------------------------
 * One linear pass over the Salome-UNV subset the reference reads (MeshReaderSalomeUnv.cpp:267-448:
 * blocks 2411 nodes, 2412 fe_id 11 boundary edges / 41 triangles -- any other fe_id is the
 * reference's "Unknown element type" error --, 2467 named groups), without its by-value string
 * lists and linear searches (:14-22, :428).  Labels are returned 0-based like the reference stores
 * them; groups are sorted by name (the reference's std::map order) and split into the cells and
 * the boundary-edge node pairs they name.  Errors: CFD2D_EINVAL + cfd2d_fvm_last_error(NULL).      */
typedef struct cfd2d_unv cfd2d_unv;
int  cfd2d_unv_read(const char* path, cfd2d_unv** out);
void cfd2d_unv_counts(const cfd2d_unv* u, int64_t* n_nodes, int64_t* n_cells, int64_t* n_bnd_edges, int32_t* n_groups);
void cfd2d_unv_copy(const cfd2d_unv* u, double* xy /*[n_nodes][2]*/, int32_t* tris /*[n_cells][3]*/, int32_t* bnd_edges /*[n_bnd_edges][2]*/);
const char* cfd2d_unv_group_name(const cfd2d_unv* u, int g);
void cfd2d_unv_group_counts(const cfd2d_unv* u, int g, int64_t* n_cells, int64_t* n_edges);
void cfd2d_unv_group_copy(const cfd2d_unv* u, int g, int64_t* cells, int32_t* edge_nodes /*[n_edges][2]*/);
void cfd2d_unv_free(cfd2d_unv* u);

const char* cfd2d_fvm_last_error(const cfd2d_fvm* h);   /* h == NULL -> last create() error         */
const char* cfd2d_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CFD2D_FVM_H */
