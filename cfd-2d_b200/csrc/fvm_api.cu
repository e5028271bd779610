// fvm_api.cu -- the C-ABI of include/cfd2d_fvm.h: handle, uploads, step orchestration (CUDA graph),
// per-kernel timing, halo exchange (NCCL, loaded at run time).  No CPU numerics live here: every
// number the caller gets back was produced by the kernels in fvm_kernels.cuh.
#include "../../include/cfd2d_fvm.h"
#include "fvm_kernels.cuh"
#include "fvm_fused.cuh"
#include "fvm_pipe.cuh"
#include "halo_nccl.h"
#include <nvtx3/nvToolsExt.h>   // header-only: ranges show up under nsys / ncu --nvtx, no-ops otherwise

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

static thread_local std::string g_create_error;
void cfd2d_set_create_error(const std::string& s) { g_create_error = s; }   // host-only helpers (unv_reader.cpp)

struct cfd2d_fvm {
    int device = 0;
    int nc = 0, nc_ex = 0, ne = 0, nmat = 0, nbc = 0;
    cfd2d_ctrl ctrl{};
    KParams P{};
    double TAU = 0.0, t = 0.0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    // device buffers
    std::vector<void*> allocs;
    double4 *Ua = nullptr, *Ub = nullptr, *W = nullptr, *Wb = nullptr, *G = nullptr, *F = nullptr;
    uint32_t* io_u32 = nullptr;   // flag staging (caller order)
    // asynchronous snapshot for FVM_TVD::save (cfd2d_fvm_snapshot_begin / _end)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_snap_done = nullptr;
    double* snap_host = nullptr;   // pinned: 5 x nc doubles (ro, ru, rv, re, cTau) + nc uint32 flags
    bool snap_pending = false;
    char* gather_buf = nullptr;    // cfd2d_fvm_gather_state staging (device)
    size_t gather_cap = 0;
    double* grad_tmp = nullptr;   // parity hook staging (cfd2d_fvm_calc_grad), allocated on first use
    // tile-fused stage kernel (fvm_fused.cuh)
    bool fused = true;
    FParams Q{};
    int ntiles = 0, n_interior = 0, n_boundary = 0, stage_nt = 256;
    size_t stage_smem = 0;
    int *d_interior = nullptr, *d_boundary = nullptr;
    // pipelined tile kernel (fvm_pipe.cuh)
    bool pipe = false;            // step layout 2: k_stage_pipe
    bool have_pipe_plan = false;
    bool w_stale = false;         // the pipe layout does not maintain the primitive cache W
    PParams PQ{};
    int pipe_nt = 256, pipe_minb = 2, pipe_grid = 0, pipe_ntiles = 0, pipe_n_interior = 0, pipe_n_boundary = 0;
    size_t pipe_smem = 0;
    int *d_pipe_interior = nullptr, *d_pipe_boundary = nullptr;
    int* d_near_send = nullptr;   // multi-rank pipe layout: owned cells within one ring of a send cell (W refresh list)
    int n_near_send = 0;
    int sm_count = 0;
    bool overlap = true;          // multi-rank: halo exchange on the comm stream, overlapped with interior work
    bool exact_riemann = false;   // Godunov: true = rim_orig_dev (the reference's operation order), false = fvm_riemann_fast.cuh
    bool lf1_cell = false;        // first-order Lax-Friedrichs: one cell-parallel sweep per stage (k_cell_lf1)
    bool diag_split = false;      // diagnostics only (CFD2D_DIAG_SPLIT=1): serial handle runs the multi-rank kernel split
    bool collective_errors = true; // multi-rank: sync() all-reduces the Newton-cap error word (CFD2D_COLLECTIVE_ERRORS=0: rank-local)
    bool skip_exchange = false;   // diagnostics only (CFD2D_DIAG_NO_EXCHANGE=1): results are wrong, timing shows the cost of the exchanges
    int ne_int = 0;               // device edges [0, ne_int) touch owned cells only; [ne_int, ne) touch a halo cell
    int *d_cells_int = nullptr, *d_cells_bnd = nullptr;   // owned cells without / with a halo neighbour
    int n_cells_int = 0, n_cells_bnd = 0;
    int* d_send_dev = nullptr;    // send cells (device ids), all peers
    int n_send = 0;
    cudaStream_t comm = nullptr;  // halo exchange stream (multi-rank handles)
    cudaEvent_t ev_G = nullptr, ev_stage = nullptr, ev_U = nullptr;
    std::vector<int> perm, orig;  // caller <-> device cell numbering
    std::unique_ptr<HostMesh> pm; // renumbered host mesh (kept for the lazily built fused plan)
    bool have_plan = false;
    std::string plan_summary;
    double* io[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // SoA staging for set/get
    unsigned long long* tau_bits = nullptr;
    int* err = nullptr;
    // graph
    bool use_graph = true;
    int graph_launches = 0;       // kernel launches inside one captured step
    int eager_steps = 0;          // steps enqueued without a graph (multi-rank: NCCL warms up eagerly first)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    // profiling
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double prof_ms[CFD2D_NKERNELS] = {0};
    int64_t prof_n[CFD2D_NKERNELS] = {0};
    int64_t launches = 0;
    // halo
    HaloNccl* halo = nullptr;
    std::vector<int> edge_pos;   // caller's edge id -> position in the device edge arrays (create())
    std::string error;
};

#define CUDA_TRY(h, call)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            char b_[512];                                                                  \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            if (h) (h)->error = b_; else g_create_error = b_;                              \
            return CFD2D_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

template <class T>
static int dev_alloc(cfd2d_fvm* h, T** p, size_t n) {
    void* q = nullptr;
    CUDA_TRY(h, cudaMalloc(&q, (n ? n : 1) * sizeof(T)));
    h->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

template <class T>
static int dev_upload(cfd2d_fvm* h, const T** p, const std::vector<T>& v) {
    T* q = nullptr;
    int rc = dev_alloc(h, &q, v.size());
    if (rc) return rc;
    if (!v.empty()) CUDA_TRY(h, cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *p = q;
    return 0;
}

// CUDA status -> CFD2D code with the message kept on the handle (create() pairs it with TRY, which
// destroys the half-built handle)
static int cuda_rc(cfd2d_fvm* h, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    h->error = std::string(what) + " failed: " + cudaGetErrorString(e);
    return CFD2D_ECUDA;
}

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

// rim_orig's constants, by the reference's own expressions (global.cpp:235-249)
static RimC make_rim(double GAM) {
    RimC k;
    k.GAM = GAM;
    k.AGAM = (GAM - 1.0);
    k.DGAM = (2.0 / k.AGAM);
    k.GGAM = (sqrt(GAM * k.AGAM));
    k.HGAM = (k.AGAM / 2.0);
    k.FGAM = (3.0 * GAM - 1.0);
    k.OGAM = (k.AGAM / (2.0 * GAM));
    k.QGAM = (GAM + 1.0);
    k.PGAM = (k.QGAM / (2.0 * GAM));
    k.RGAM = (4.0 * GAM);
    k.SGAM = (GAM * k.AGAM);
    k.TGAM = (k.QGAM / 2.0);
    k.IAGAM = (1 / k.AGAM);
    k.DG1 = (1 + k.DGAM);
    k.DGGG = k.DGAM * k.GGAM;
    k.ISGAM = 1.0 / k.SGAM;
    k.IDG1 = 1.0 / k.DG1;
    return k;
}

// ---- kernel launch helpers with optional per-kernel event timing ---------------------------
struct NvtxScope { NvtxScope(const char* n) { nvtxRangePushA(n); } ~NvtxScope() { nvtxRangePop(); } };
static const char* const k_names[CFD2D_NKERNELS] = {"k_grad", "k_flux", "k_update<1>", "k_update<2>", "k_remediate", "k_tau_steady",
                                                     "halo exchange", "stage 1 (fused)", "stage 2 (fused)"};
struct KTimer {
    cfd2d_fvm* h; int id; cudaStream_t st;
    KTimer(cfd2d_fvm* h_, int id_, cudaStream_t st_ = nullptr) : h(h_), id(id_), st(st_ ? st_ : h_->stream) {
        nvtxRangePushA(k_names[id]);
        if (h->profiling) cudaEventRecord(h->ev0, st);
    }
    ~KTimer() {
        nvtxRangePop();
        h->launches++;
        if (h->profiling) {
            cudaEventRecord(h->ev1, st);
            cudaEventSynchronize(h->ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->ev0, h->ev1);
            h->prof_ms[id] += ms;
            h->prof_n[id]++;
        }
    }
};

static void launch_prim(cfd2d_fvm* h, const double4* U, double4* W, int c0, int c1, cudaStream_t st) {
    if (c1 <= c0) return;
    h->launches++;
    k_prim<<<nblk(c1 - c0, 256), 256, 0, st>>>(h->P, U, W, c0, c1);
}

// list == nullptr: all owned cells
// interior: all owned cells but those with a rank-halo neighbour (the boundary pass's list)
static void launch_grad(cfd2d_fvm* h, const int* list = nullptr, int n = -1, cudaStream_t st = nullptr, bool interior = false) {
    if (!list && n < 0) n = h->nc;
    if (n <= 0) return;
    if (!st) st = h->stream;
    KTimer t(h, CFD2D_K_GRAD, st);
    k_grad<<<nblk(n, CFD2D_GRAD_NT), CFD2D_GRAD_NT, 0, st>>>(h->P, h->W, h->G, list, n, interior ? 1 : 0);
}

// device edges [e0, e1); e1 < 0: all edges
static void launch_flux(cfd2d_fvm* h, const double4* Ucur, int scale, int e0 = 0, int e1 = -1, cudaStream_t st = nullptr) {
    if (e1 < 0) e1 = h->ne;
    if (e1 <= e0) return;
    if (!st) st = h->stream;
    KTimer t(h, CFD2D_K_FLUX, st);
    dim3 g(nblk(2 * (long long)(e1 - e0), CFD2D_FLUX_NT)), b(CFD2D_FLUX_NT);   // one thread per (edge, Gauss point)
    const int fv = (h->ctrl.flux == CFD2D_FLUX_GODUNOV) ? (h->exact_riemann ? 0 : 2) : 1, od = h->ctrl.order;
    typedef void (*flux_fn)(KParams, const double4*, const double4*, const double4*, double4*, int, int, int);
    flux_fn f;
    f = fv == 2 ? (od == 2 ? k_flux<2, 2> : k_flux<2, 1>)
      : fv == 0 ? (od == 2 ? k_flux<0, 2> : k_flux<0, 1>)
                : (od == 2 ? k_flux<1, 2> : k_flux<1, 1>);
    f<<<g, b, 0, st>>>(h->P, h->W, h->G, Ucur, h->F, scale, e0, e1);
}

static void launch_update(cfd2d_fvm* h, int stage) {
    if (h->nc == 0) return;
    KTimer t(h, stage == 1 ? CFD2D_K_UPDATE1 : CFD2D_K_UPDATE2);
    if (stage == 1) k_update<1><<<nblk(h->nc, CFD2D_UPDATE_NT), CFD2D_UPDATE_NT, 0, h->stream>>>(h->P, h->F, h->Ua, h->Ub, h->W);
    else k_update<2><<<nblk(h->nc, CFD2D_UPDATE_NT), CFD2D_UPDATE_NT, 0, h->stream>>>(h->P, h->F, h->Ub, h->Ua, h->W);
}

static void launch_remediate(cfd2d_fvm* h) {
    KTimer t(h, CFD2D_K_REMEDIATE);
    // Ub (the stage-1 state) and G are dead at the end of a step: snapshot / staging space of the sweep
    k_remediate<<<1, 1024, 0, h->stream>>>(h->P, h->Ua, h->Ub, h->W, h->G);
}

static void launch_tau_steady(cfd2d_fvm* h) {
    if (h->nc == 0) return;
    KTimer t(h, CFD2D_K_TIMESTEP);
    k_tau_steady<<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->P, h->W);
}

// ---- the tile-fused stage kernel: template dispatch over (flux, order, stage, block size) ------
typedef void (*stage_fn)(KParams, FParams, const double4*, const double4*, double4*, double4*, const double4*);

template <int NT, int MINB>
static stage_fn stage_kernel_nt(int flux, int order, int stage) {
    if (flux == CFD2D_FLUX_GODUNOV) {
        if (order == 2) return stage == 1 ? (stage_fn)k_stage<0, 2, 1, NT, MINB> : (stage_fn)k_stage<0, 2, 2, NT, MINB>;
        return stage == 1 ? (stage_fn)k_stage<0, 1, 1, NT, MINB> : (stage_fn)k_stage<0, 1, 2, NT, MINB>;
    }
    if (flux == 2) {    // Godunov through the reduced-instruction solver
        if (order == 2) return stage == 1 ? (stage_fn)k_stage<2, 2, 1, NT, MINB> : (stage_fn)k_stage<2, 2, 2, NT, MINB>;
        return stage == 1 ? (stage_fn)k_stage<2, 1, 1, NT, MINB> : (stage_fn)k_stage<2, 1, 2, NT, MINB>;
    }
    if (order == 2) return stage == 1 ? (stage_fn)k_stage<1, 2, 1, NT, MINB> : (stage_fn)k_stage<1, 2, 2, NT, MINB>;
    return stage == 1 ? (stage_fn)k_stage<1, 1, 1, NT, MINB> : (stage_fn)k_stage<1, 1, 2, NT, MINB>;
}

static int flux_variant(const cfd2d_fvm* h) {   // template FLUX value: 0 Godunov bit-faithful, 1 LF, 2 Godunov reduced-instruction
    return (h->ctrl.flux == CFD2D_FLUX_GODUNOV && !h->exact_riemann) ? 2 : h->ctrl.flux;
}

static stage_fn stage_kernel(const cfd2d_fvm* h, int stage) {
    const int fx = flux_variant(h);
    switch (h->stage_nt) {
        case 128: return stage_kernel_nt<128, 8>(fx, h->ctrl.order, stage);
        case 384: return stage_kernel_nt<384, 2>(fx, h->ctrl.order, stage);
        case 512: return stage_kernel_nt<512, 2>(fx, h->ctrl.order, stage);
        default:  return stage_kernel_nt<256, 4>(fx, h->ctrl.order, stage);
    }
}

// tiles: nullptr = all tiles [0, ntiles), else a list of n tile ids
static void launch_stage(cfd2d_fvm* h, int stage, const int* tiles, int n) {
    if (n <= 0) return;
    KTimer t(h, stage == 1 ? CFD2D_K_STAGE1 : CFD2D_K_STAGE2);
    FParams q = h->Q;
    q.tile_ids = tiles;
    stage_fn f = stage_kernel(h, stage);
    if (stage == 1) f<<<n, h->stage_nt, h->stage_smem, h->stream>>>(h->P, q, h->W, h->Ua, h->Ub, h->Wb, h->G);
    else            f<<<n, h->stage_nt, h->stage_smem, h->stream>>>(h->P, q, h->Wb, h->Ub, h->Ua, h->W, h->G);
}

// halo exchange of U4 (rec4 = 1) into `U`'s halo slice followed by the halo prim conversion into
// `W`, or of G8 (rec4 = 2).  Method::exchange semantics (method.h:13-41): recv lands contiguously at
// recvShift.
static int exchange_U(cfd2d_fvm* h, double4* U, double4* W, cudaStream_t st) {
    if (!h->halo) return 0;
    KTimer t(h, CFD2D_K_HALO, st);
    int rc = h->skip_exchange ? 0 : halo_exchange(h->halo, U, 1, st, &h->launches);
    if (rc) { h->error = halo_error(h->halo); return rc; }
    launch_prim(h, U, W, h->nc, h->nc_ex, st);
    return 0;
}

static int exchange_G(cfd2d_fvm* h, cudaStream_t st) {
    if (!h->halo || h->ctrl.order != 2) return 0;
    KTimer t(h, CFD2D_K_HALO, st);
    int rc = h->skip_exchange ? 0 : halo_exchange(h->halo, h->G, 2, st, &h->launches);
    if (rc) h->error = halo_error(h->halo);
    return rc;
}

// One whole RK2 step = the body of the while loop of FVM_TVD::run (fvm_tvd.cpp:310-450), three
// sweeps per stage (the layout of the reference).
//
// Multi-rank handles (two neighbour exchanges per stage).  Everything that touches halo
// data -- the exchanges AND the small partition-boundary kernels -- runs on the high-priority comm
// stream, concurrently with the interior sweeps on the compute stream:
//   comm   : [exchange U of stage 1] grad(cells with a halo neighbour) | exchange G | flux(edges touching a halo cell)
//   compute: grad(other cells) ......................................^ flux(interior edges) ............^ update
// so the critical path of a stage is the same three sweeps as on one GPU.  The state exchange after
// stage 2 is the only exposed one (remediateLimCells reads halo states).
static int enqueue_step_unfused(cfd2d_fvm* h) {
    int rc;
    const bool multi = h->halo != nullptr;
    const bool ov = multi && h->overlap;
    cudaStream_t S = h->stream, C = ov ? h->comm : h->stream;
    const bool o2 = h->ctrl.order == 2;
    if (h->ctrl.steady) launch_tau_steady(h);                 // :315
    if (h->lf1_cell) {
        // first-order LF: no gradients, fluxes evaluated from both sides => the stage is ONE sweep.
        // Multi-rank: the cells with a halo neighbour run on the comm stream after the state exchange,
        // all others on the compute stream meanwhile.  W ping-pongs: stage 1 W -> Wb, stage 2 Wb -> W.
        for (int stage = 1; stage <= 2; stage++) {
            const double4* Wc = stage == 1 ? h->W : h->Wb;
            double4* Wo = stage == 1 ? h->Wb : h->W;
            double4* Ui = stage == 1 ? h->Ua : h->Ub;
            double4* Uo = stage == 1 ? h->Ub : h->Ua;
            const int kid = stage == 1 ? CFD2D_K_STAGE1 : CFD2D_K_STAGE2;
            auto sweep = [&](const int* list, int n, cudaStream_t st) {
                if (n <= 0) return;
                KTimer t(h, kid, st);
                if (stage == 1) k_cell_lf1<1><<<nblk(n, 256), 256, 0, st>>>(h->P, Wc, Ui, Uo, Wo, list, n);
                else            k_cell_lf1<2><<<nblk(n, 256), 256, 0, st>>>(h->P, Wc, Ui, Uo, Wo, list, n);
            };
            if (!multi) { sweep(nullptr, h->nc, S); continue; }
            if (ov) { cudaEventRecord(h->ev_stage, S); cudaStreamWaitEvent(C, h->ev_stage, 0); }   // fork
            // halo copy of the state this stage starts from: stage 2 = the stage-1 result; stage 1 = Ua again,
            // because remediateLimCells may have rewritten send cells after the end-of-step exchange
            if ((rc = exchange_U(h, Ui, stage == 1 ? h->W : h->Wb, C))) return rc;
            sweep(h->d_cells_bnd, h->n_cells_bnd, C);
            if (ov) cudaEventRecord(h->ev_U, C);
            sweep(h->d_cells_int, h->n_cells_int, S);
            if (ov) cudaStreamWaitEvent(S, h->ev_U, 0);                               // join
        }
        if (multi) {
            if (ov) { cudaEventRecord(h->ev_stage, S); cudaStreamWaitEvent(C, h->ev_stage, 0); }
            if ((rc = exchange_U(h, h->Ua, h->W, C))) return rc;
            if (ov) { cudaEventRecord(h->ev_U, C); cudaStreamWaitEvent(S, h->ev_U, 0); }
        }
        launch_remediate(h);
        return 0;
    }
    for (int stage = 1; stage <= 2; stage++) {
        double4* Ucur = stage == 1 ? h->Ua : h->Ub;          // state this stage starts from
        if (!multi && !h->diag_split) {
            if (o2) launch_grad(h);
            launch_flux(h, Ucur, 1);
        } else {
            if (ov) { cudaEventRecord(h->ev_stage, S); cudaStreamWaitEvent(C, h->ev_stage, 0); }   // fork
            // halo copy of the state this stage starts from (hidden behind the interior gradient sweep).
            // Stage 1 re-sends Ua: remediateLimCells (end of the previous step) may have rewritten send
            // cells AFTER the end-of-step exchange, and the peers' stage-1 gradients/fluxes must see them
            // (the serial reference sweeps every cell before the next step, fvm_tvd.cpp:449).
            if ((rc = exchange_U(h, Ucur, h->W, C))) return rc;
            if (o2) {
                launch_grad(h, h->d_cells_bnd, h->n_cells_bnd, C);
                if (ov) cudaEventRecord(h->ev_G, C);
                if ((rc = exchange_G(h, C))) return rc;
                launch_grad(h, nullptr, -1, S, !h->diag_split || multi);   // diag_split on a serial handle keeps the list form
                if (ov) cudaStreamWaitEvent(S, h->ev_G, 0);   // interior edges read the gradients of all owned cells
            }
            launch_flux(h, Ucur, 1, h->ne_int, h->ne, C);
            if (ov) cudaEventRecord(h->ev_U, C);
            launch_flux(h, Ucur, 1, 0, h->ne_int, S);
            if (ov) cudaStreamWaitEvent(S, h->ev_U, 0);       // join
        }
        launch_update(h, stage);                               // stage 1: Ub, W; stage 2: Ua, W, flags (:366-374, :419-447)
    }
    if (multi) {
        if (ov) { cudaEventRecord(h->ev_stage, S); cudaStreamWaitEvent(C, h->ev_stage, 0); }
        if ((rc = exchange_U(h, h->Ua, h->W, C))) return rc;
        if (ov) { cudaEventRecord(h->ev_U, C); cudaStreamWaitEvent(S, h->ev_U, 0); }
    }
    launch_remediate(h);                                       // :449 (no-op kernel when nothing is flagged)
    return 0;
}

// The same step with one fused kernel per stage.  Serial handle: 3 launches per step.
// Multi-rank handle: per stage, the comm stream computes the gradients of the cells a peer needs
// (k_grad on the send list), exchanges them, and later exchanges the new state, while the compute
// stream runs the INTERIOR tiles (no rank-halo data within two rings); the BOUNDARY tiles wait for
// the gradient exchange.  W is ping-ponged: stage 1 reads W writes Wb, stage 2 reads Wb writes W.
static int enqueue_step_fused(cfd2d_fvm* h) {
    int rc;
    if (h->ctrl.steady) launch_tau_steady(h);
    if (!h->halo) {
        launch_stage(h, 1, nullptr, h->ntiles);
        launch_stage(h, 2, nullptr, h->ntiles);
        launch_remediate(h);
        return 0;
    }
    for (int stage = 1; stage <= 2; stage++) {
        const double4* Wcur = stage == 1 ? h->W : h->Wb;
        double4* Uout = stage == 1 ? h->Ub : h->Ua;
        double4* Wout = stage == 1 ? h->Wb : h->W;
        // comm stream: everything before this point on the compute stream (previous stage, its
        // exchange) is complete
        cudaEventRecord(h->ev_stage, h->stream);
        cudaStreamWaitEvent(h->comm, h->ev_stage, 0);
        // cells remediated at the end of the previous step -> peers (see enqueue_step_unfused)
        if (stage == 1 && (rc = exchange_U(h, h->Ua, h->W, h->comm))) return rc;
        if (h->ctrl.order == 2) {
            if (h->n_send > 0) {
                h->launches++;
                k_grad<<<nblk(h->n_send, CFD2D_GRAD_NT), CFD2D_GRAD_NT, 0, h->comm>>>(h->P, Wcur, h->G, h->d_send_dev, h->n_send, 0);
            }
            if ((rc = exchange_G(h, h->comm))) return rc;
        }
        cudaEventRecord(h->ev_G, h->comm);
        launch_stage(h, stage, h->d_interior, h->n_interior);
        cudaStreamWaitEvent(h->stream, h->ev_G, 0);
        launch_stage(h, stage, h->d_boundary, h->n_boundary);
        // new state of the send cells -> peers; the next stage's interior tiles do not need it
        cudaEventRecord(h->ev_stage, h->stream);
        cudaStreamWaitEvent(h->comm, h->ev_stage, 0);
        if ((rc = exchange_U(h, Uout, Wout, h->comm))) return rc;
        cudaEventRecord(h->ev_U, h->comm);
        if (stage == 2) {
            // remediateLimCells reads neighbour states, possibly halo cells: after the exchange
            cudaStreamWaitEvent(h->stream, h->ev_U, 0);
            launch_remediate(h);
        }
    }
    // the next step's comm work is ordered after ev_U by the comm stream itself; its boundary tiles
    // wait on ev_G, recorded after it
    return 0;
}

// ---- the pipelined tile kernel: template dispatch over (flux, order, stage, block size) ---------
typedef void (*pipe_fn)(KParams, PParams, const double4*, double4*, const double4*);

template <int NT, int MINB>
static pipe_fn pipe_kernel_nt(int flux, int order, int stage) {
    if (flux == CFD2D_FLUX_GODUNOV) {
        if (order == 2) return stage == 1 ? (pipe_fn)k_stage_pipe<0, 2, 1, NT, MINB> : (pipe_fn)k_stage_pipe<0, 2, 2, NT, MINB>;
        return stage == 1 ? (pipe_fn)k_stage_pipe<0, 1, 1, NT, MINB> : (pipe_fn)k_stage_pipe<0, 1, 2, NT, MINB>;
    }
    if (flux == 2) {
        if (order == 2) return stage == 1 ? (pipe_fn)k_stage_pipe<2, 2, 1, NT, MINB> : (pipe_fn)k_stage_pipe<2, 2, 2, NT, MINB>;
        return stage == 1 ? (pipe_fn)k_stage_pipe<2, 1, 1, NT, MINB> : (pipe_fn)k_stage_pipe<2, 1, 2, NT, MINB>;
    }
    if (order == 2) return stage == 1 ? (pipe_fn)k_stage_pipe<1, 2, 1, NT, MINB> : (pipe_fn)k_stage_pipe<1, 2, 2, NT, MINB>;
    return stage == 1 ? (pipe_fn)k_stage_pipe<1, 1, 1, NT, MINB> : (pipe_fn)k_stage_pipe<1, 1, 2, NT, MINB>;
}

// (threads per CTA, min resident CTAs per SM) pairs that are compiled: the second number caps the
// registers (65536 / (NT * MINB)), the shared memory of the tile decides what is really resident
static pipe_fn pipe_kernel(const cfd2d_fvm* h, int stage) {
    const int fx = flux_variant(h), od = h->ctrl.order;
    switch (h->pipe_nt * 10 + h->pipe_minb) {
#ifndef CFD2D_PIPE_FEW      /* kernel-variant sweeps: build only the default shape */
        case 1284:  return pipe_kernel_nt<128, 4>(fx, od, stage);
        case 2563:  return pipe_kernel_nt<256, 3>(fx, od, stage);
        case 3841:  return pipe_kernel_nt<384, 1>(fx, od, stage);
        case 3842:  return pipe_kernel_nt<384, 2>(fx, od, stage);
        case 5121:  return pipe_kernel_nt<512, 1>(fx, od, stage);
        case 5122:  return pipe_kernel_nt<512, 2>(fx, od, stage);
        case 7681:  return pipe_kernel_nt<768, 1>(fx, od, stage);
        case 10241: return pipe_kernel_nt<1024, 1>(fx, od, stage);
#endif
        default:    return pipe_kernel_nt<256, 2>(fx, od, stage);
    }
}

// tiles: nullptr = all tiles, else a list of n tile ids.  Persistent launch: at most one wave of CTAs.
static void launch_pipe(cfd2d_fvm* h, int stage, const int* tiles, int n, cudaStream_t st = nullptr) {
    if (n <= 0) return;
    if (!st) st = h->stream;
    KTimer t(h, stage == 1 ? CFD2D_K_STAGE1 : CFD2D_K_STAGE2, st);
    PParams q = h->PQ;
    q.tile_ids = tiles;
    q.n_tiles = n;
    const int grid = n < h->pipe_grid ? n : h->pipe_grid;
    pipe_fn f = pipe_kernel(h, stage);
    if (stage == 1) f<<<grid, h->pipe_nt, h->pipe_smem, st>>>(h->P, q, h->Ua, h->Ub, h->G);
    else            f<<<grid, h->pipe_nt, h->pipe_smem, st>>>(h->P, q, h->Ub, h->Ua, h->G);
}

// The step with the pipelined tile kernel: 3 launches (k_stage_pipe x 2, k_remediate); W is not
// maintained (w_stale).  Multi-rank handles: per stage the comm stream refreshes the state halo,
// converts the cells around the send set to primitive form, computes and exchanges the gradients of
// the send cells; interior tiles (no rank-halo data within two rings) run meanwhile, boundary tiles
// after the gradient exchange.
static int enqueue_step_pipe(cfd2d_fvm* h) {
    int rc;
    if (h->ctrl.steady) {                                    // calcTimeStep reads the primitive state (:315)
        launch_prim(h, h->Ua, h->W, 0, h->nc, h->stream);
        launch_tau_steady(h);
    }
    if (!h->halo) {
        launch_pipe(h, 1, nullptr, h->pipe_ntiles);
        launch_pipe(h, 2, nullptr, h->pipe_ntiles);
        launch_remediate(h);
        return 0;
    }
    for (int stage = 1; stage <= 2; stage++) {
        double4* Ucur = stage == 1 ? h->Ua : h->Ub;
        cudaEventRecord(h->ev_stage, h->stream);
        cudaStreamWaitEvent(h->comm, h->ev_stage, 0);
        if ((rc = exchange_U(h, Ucur, h->W, h->comm))) return rc;        // halo U (+ halo W for the send cells' gradients)
        if (h->ctrl.order == 2) {
            if (h->n_near_send > 0) {
                h->launches++;
                k_prim_list<<<nblk(h->n_near_send, 256), 256, 0, h->comm>>>(h->P, Ucur, h->W, h->d_near_send, h->n_near_send);
            }
            if (h->n_send > 0) {
                h->launches++;
                k_grad<<<nblk(h->n_send, CFD2D_GRAD_NT), CFD2D_GRAD_NT, 0, h->comm>>>(h->P, h->W, h->G, h->d_send_dev, h->n_send, 0);
            }
            if ((rc = exchange_G(h, h->comm))) return rc;
        }
        cudaEventRecord(h->ev_G, h->comm);
        launch_pipe(h, stage, h->d_pipe_interior, h->pipe_n_interior);
        cudaStreamWaitEvent(h->stream, h->ev_G, 0);
        launch_pipe(h, stage, h->d_pipe_boundary, h->pipe_n_boundary);
    }
    // remediateLimCells reads neighbour states, possibly halo cells: the end-of-step exchange first
    cudaEventRecord(h->ev_stage, h->stream);
    cudaStreamWaitEvent(h->comm, h->ev_stage, 0);
    if ((rc = exchange_U(h, h->Ua, h->W, h->comm))) return rc;
    cudaEventRecord(h->ev_U, h->comm);
    cudaStreamWaitEvent(h->stream, h->ev_U, 0);
    launch_remediate(h);
    return 0;
}

static int enqueue_step(cfd2d_fvm* h) {
    if (h->pipe) { h->w_stale = true; return enqueue_step_pipe(h); }
    return h->fused ? enqueue_step_fused(h) : enqueue_step_unfused(h);
}

// consumers of the primitive cache outside the step (time step, parity hooks, layout switches)
static void ensure_W(cfd2d_fvm* h) {
    if (!h->w_stale) return;
    launch_prim(h, h->Ua, h->W, 0, h->nc_ex, h->stream);
    h->w_stale = false;
}

static void drop_graph(cfd2d_fvm* h) {
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
}

static int check_device_errors(cfd2d_fvm* h) {
    int e[2] = {0, 0};
    CUDA_TRY(h, cudaMemcpyAsync(e, h->err, sizeof e, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->halo && h->collective_errors) {
        // a rank that stops stepping would leave its peers blocked in the next ncclSend/Recv: every rank
        // learns the worst count (MPI_Allreduce analogue, fem_rkdg.cpp:413) and returns the same code
        double worst = -(double)e[0];
        int rc = halo_allreduce_min(h->halo, &worst, h->comm);
        if (rc) { h->error = halo_error(h->halo); return rc; }
        if (worst < 0.0 && e[0] == 0) e[0] = (int)(-worst);
    }
    if (e[0] != 0) {
        char b[256];
        snprintf(b, sizeof b, "rim_orig Newton iteration hit the cap (%d) on %d edge evaluations: non-physical "
                 "left/right states (the reference would loop forever here)", h->P.max_newton, e[0]);
        h->error = b;
        cudaMemsetAsync(h->err, 0, sizeof(int), h->stream);
        return CFD2D_ENEWTON;
    }
    return 0;
}

// Host plan + device tables of the tile-fused stage kernel (fvm_tiling.h, fvm_fused.cuh); built on
// first use from the renumbered mesh the handle keeps.
static int build_fused_plan(cfd2d_fvm* h) {
    if (h->have_plan) return 0;
    if (!h->pm) { h->error = "internal: host mesh released"; return CFD2D_EINVAL; }
    const HostMesh& pm = *h->pm;
    const int nc = h->nc;
    int rc = 0;
#define FTRY(x) do { rc = (x); if (rc) return rc; } while (0)
    {
        int TC = 512;
        if (const char* ev = getenv("CFD2D_TILE")) TC = atoi(ev);
        h->stage_nt = 512;
        if (const char* ev = getenv("CFD2D_NT")) h->stage_nt = atoi(ev);
        if (h->stage_nt != 128 && h->stage_nt != 256 && h->stage_nt != 384 && h->stage_nt != 512) h->stage_nt = 256;
        TilePlan tp;
        std::string terr = build_tile_plan(pm, TC, tp);
        if (!terr.empty()) { h->error = terr; return CFD2D_EINVAL; }
        h->ntiles = tp.ntiles;
        h->n_interior = (int)tp.interior.size();
        h->n_boundary = (int)tp.boundary.size();
        const size_t net = tp.e_c1.size();
        std::vector<double2> t_n(net);
        std::vector<double> t_l2(net);
        std::vector<double4> t_d1(net), t_d2(net);
        for (size_t q = 0; q < net; q++) {
            const int e = tp.e_id[q];
            const int c1 = pm.edge_c1[e], c2 = pm.edge_c2[e];
            t_n[q] = make_double2(pm.edge_nx[e], pm.edge_ny[e]);
            t_l2[q] = pm.edge_l[e] * 0.5;                               // fvm_tvd.cpp:335
            const double* g = pm.edge_gp.data() + 4 * (size_t)e;
            t_d1[q] = make_double4(g[0] - pm.cell_cx[c1], g[1] - pm.cell_cy[c1], g[2] - pm.cell_cx[c1], g[3] - pm.cell_cy[c1]);
            if (c2 >= 0) t_d2[q] = make_double4(g[0] - pm.cell_cx[c2], g[1] - pm.cell_cy[c2], g[2] - pm.cell_cx[c2], g[3] - pm.cell_cy[c2]);
            else t_d2[q] = make_double4(0, 0, 0, 0);
        }
        FParams& Q = h->Q;
        Q.tile_ids = nullptr;
        Q.nl_max = tp.nl_max; Q.ne_max = tp.ne_max;
        FTRY(dev_upload(h, &Q.tiles, tp.tiles));
        FTRY(dev_upload(h, &Q.ring, tp.ring));
        FTRY(dev_upload(h, &Q.g_nb, tp.g_nb));
        FTRY(dev_upload(h, &Q.g_nx, tp.g_nx));
        FTRY(dev_upload(h, &Q.g_ny, tp.g_ny));
        FTRY(dev_upload(h, &Q.g_l, tp.g_l));
        FTRY(dev_upload(h, &Q.e_c1, tp.e_c1));
        FTRY(dev_upload(h, &Q.e_c2, tp.e_c2));
        FTRY(dev_upload(h, &Q.e_cl, tp.e_cl));
        FTRY(dev_upload(h, &Q.e_n, t_n));
        FTRY(dev_upload(h, &Q.e_l2, t_l2));
        FTRY(dev_upload(h, &Q.e_d1, t_d1));
        FTRY(dev_upload(h, &Q.e_d2, t_d2));
        FTRY(dev_upload(h, &Q.u_es, tp.u_es));
        Q.c_orig = h->P.c_orig;
        { const int* q = nullptr; FTRY(dev_upload(h, &q, tp.interior)); h->d_interior = (int*)q; }
        { const int* q = nullptr; FTRY(dev_upload(h, &q, tp.boundary)); h->d_boundary = (int*)q; }
        h->stage_smem = ((h->ctrl.order == 2 ? 6 * (size_t)tp.nl_max : 2 * (size_t)tp.nl_max) + 2 * (size_t)tp.ne_max) * sizeof(double2)
                        + (h->ctrl.flux == CFD2D_FLUX_LAX ? (size_t)tp.nl_max * sizeof(double) : 0);
        if (h->stage_smem > 227 * 1024) { h->error = "tile does not fit in shared memory (lower CFD2D_TILE)"; return CFD2D_EINVAL; }
        for (int st = 1; st <= 2; st++) {
            stage_fn f = stage_kernel(h, st);
            CUDA_TRY(h, cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->stage_smem));
        }
        char b[256];
        snprintf(b, sizeof b, "tiles=%d (TC=%d, interior=%d, boundary=%d) nl_max=%d ne_max=%d smem=%zu B ring/own=%.3f edges/own=%.3f nt=%d",
                 tp.ntiles, tp.TC, h->n_interior, h->n_boundary, tp.nl_max, tp.ne_max, h->stage_smem,
                 nc ? (double)tp.sum_ring / nc : 0.0, nc ? (double)tp.sum_ne / nc : 0.0, h->stage_nt);
        h->plan_summary = b;
    }
#undef FTRY
    h->have_plan = true;
    return 0;
}

// Host plan + device tables of the pipelined tile kernel (fvm_tiling.h build_pipe_plan, fvm_pipe.cuh)
static int build_pipe_plan_dev(cfd2d_fvm* h) {
    if (h->have_pipe_plan) return 0;
    if (!h->pm) { h->error = "internal: host mesh released"; return CFD2D_EINVAL; }
    int rc = 0;
#define FTRY(x) do { rc = (x); if (rc) return rc; } while (0)
    int TC = 128;
    if (const char* ev = getenv("CFD2D_PIPE_TILE")) TC = atoi(ev);
    h->pipe_nt = 256;
    if (const char* ev = getenv("CFD2D_PIPE_NT")) h->pipe_nt = atoi(ev);
    h->pipe_minb = 0;
    if (const char* ev = getenv("CFD2D_PIPE_MINB")) h->pipe_minb = atoi(ev);
    {
        static const int ok[][2] = {{128, 4}, {256, 2}, {256, 3}, {384, 1}, {384, 2}, {512, 1}, {512, 2}, {768, 1}, {1024, 1}};
        bool found = false, nt_known = false;
        int first_minb = 0;
        for (auto& o : ok) {
            if (o[0] == h->pipe_nt) { if (!nt_known) first_minb = o[1]; nt_known = true; if (o[1] == h->pipe_minb) found = true; }
        }
        if (!nt_known) { h->pipe_nt = 256; h->pipe_minb = 2; }
        else if (!found) h->pipe_minb = first_minb;
    }
    PipePlan pp;
    // Godunov: edges of a tile grouped by normal direction (branch coherence of rim_orig); LF: by cell id
    std::string perr = build_pipe_plan(*h->pm, TC, h->ctrl.flux == CFD2D_FLUX_GODUNOV, pp);
    if (!perr.empty()) { h->error = perr; return CFD2D_EINVAL; }
    h->pipe_ntiles = pp.ntiles;
    h->pipe_n_interior = (int)pp.interior.size();
    h->pipe_n_boundary = (int)pp.boundary.size();
    PParams& Q = h->PQ;
    Q.tile_ids = nullptr; Q.n_tiles = pp.ntiles;
    FTRY(dev_upload(h, &Q.tiles, pp.tiles));
    FTRY(dev_upload(h, &Q.ring, pp.ring));
    FTRY(dev_upload(h, &Q.blob, pp.blob));
    Q.c_orig = h->P.c_orig;
    { const int* q = nullptr; FTRY(dev_upload(h, &q, pp.interior)); h->d_pipe_interior = (int*)q; }
    { const int* q = nullptr; FTRY(dev_upload(h, &q, pp.boundary)); h->d_pipe_boundary = (int*)q; }
    auto up = [](long long x, long long a) { return (int)((x + a - 1) / a * a); };
    const int TCp = pp.TC;
    int o = 0;
    o = up(pp.blob_max, 128);
    Q.so_u = o; o += 32 * TCp;
    Q.so_uold = o; o += 32 * TCp;
    Q.so_cfl = o; o += up(8LL * TCp + 16, 16);
    Q.so_flag = o; o += up(4LL * TCp + 16, 16);
    Q.so_ring = o; o += 32 * (pp.nring_max > 0 ? pp.nring_max : 1);
    Q.so_gx = o; o += 64 * (pp.nhalo_max > 0 ? pp.nhalo_max : 1);
    Q.stage_bytes = up(o, 128);
    Q.o_stage0 = PIPE_HDR_BYTES;
    int q = Q.o_stage0 + 2 * Q.stage_bytes;
    Q.nl2_max = pp.nl2_max; Q.nl_max = pp.nl_max; Q.ne_max = pp.ne_max;
    Q.o_W0 = q; q += 16 * pp.nl2_max;
    Q.o_W1 = q; q += 16 * pp.nl2_max;
    Q.o_E = q; q += up(8LL * (h->ctrl.flux == CFD2D_FLUX_LAX ? pp.nl2_max : 0), 16);
    Q.o_G = q; q += (h->ctrl.order == 2 ? 64 * pp.nl_max : 0);
    Q.o_F = q; q += 32 * pp.ne_max;
    h->pipe_smem = (size_t)q;
    if (h->pipe_smem > 227 * 1024) { h->error = "pipe tile does not fit in shared memory (lower CFD2D_PIPE_TILE)"; return CFD2D_EINVAL; }
    for (int fx = 0; fx < 2; fx++) {        // both Riemann variants of a Godunov handle
        const bool keep = h->exact_riemann;
        h->exact_riemann = fx != 0;
        for (int st = 1; st <= 2; st++)
            CUDA_TRY(h, cudaFuncSetAttribute((const void*)pipe_kernel(h, st), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->pipe_smem));
        h->exact_riemann = keep;
    }
    {
        cudaDeviceProp prop;
        CUDA_TRY(h, cudaGetDeviceProperties(&prop, h->device));
        h->sm_count = prop.multiProcessorCount;
        int per_sm = 0;
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)pipe_kernel(h, 2), h->pipe_nt, h->pipe_smem));
        if (per_sm < 1) { h->error = "k_stage_pipe cannot be resident (registers / shared memory)"; return CFD2D_EINVAL; }
        if (const char* ev = getenv("CFD2D_PIPE_CTAS")) { int v = atoi(ev); if (v > 0 && v < per_sm) per_sm = v; }
        h->pipe_grid = h->sm_count * per_sm;          // persistent: one resident wave
        char b[320];
        snprintf(b, sizeof b, "pipe: tiles=%d (TC=%d, interior=%d, boundary=%d) nt=%d ctas/sm=%d grid=%d smem=%zu B (stage %d B, blob<=%d B) "
                 "ring1/own=%.3f ring2/own=%.3f edges/own=%.3f blob B/cell=%.1f",
                 pp.ntiles, pp.TC, h->pipe_n_interior, h->pipe_n_boundary, h->pipe_nt, per_sm, h->pipe_grid, h->pipe_smem, Q.stage_bytes, pp.blob_max,
                 h->nc ? (double)pp.sum_ring1 / h->nc : 0.0, h->nc ? (double)pp.sum_ring2 / h->nc : 0.0, h->nc ? (double)pp.sum_ne / h->nc : 0.0,
                 h->nc ? (double)pp.blob.size() / h->nc : 0.0);
        h->plan_summary = b;
    }
    if (h->halo) {
        // owned cells within one ring of a send cell: their W feeds k_grad(send list)
        const HostMesh& pm = *h->pm;
        std::vector<char> mark(h->nc, 0);
        std::vector<int> send_dev(h->n_send > 0 ? h->n_send : 0);
        if (h->n_send > 0) CUDA_TRY(h, cudaMemcpy(send_dev.data(), h->d_send_dev, (size_t)h->n_send * sizeof(int), cudaMemcpyDeviceToHost));
        for (int c : send_dev) {
            mark[c] = 1;
            for (int k = 0; k < 3; k++) {
                const int e = pm.cell_edges[3 * (size_t)c + k];
                const int nb = pm.edge_c1[e] == c ? pm.edge_c2[e] : pm.edge_c1[e];
                if (nb >= 0 && nb < h->nc) mark[nb] = 1;
            }
        }
        std::vector<int> near;
        for (int c = 0; c < h->nc; c++) if (mark[c]) near.push_back(c);
        h->n_near_send = (int)near.size();
        { const int* q2 = nullptr; FTRY(dev_upload(h, &q2, near)); h->d_near_send = (int*)q2; }
    }
#undef FTRY
    h->have_pipe_plan = true;
    return 0;
}

extern "C" {

const char* cfd2d_version(void) { return "cfd2d_b200 0.1 (sm_100a)"; }

const char* cfd2d_fvm_last_error(const cfd2d_fvm* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int cfd2d_fvm_create(const cfd2d_mesh* m, const cfd2d_phys* p, const cfd2d_ctrl* c, const cfd2d_halo* halo,
                     int device, cfd2d_fvm** out) {
    g_create_error.clear();
    if (!m || !p || !c || !out) { g_create_error = "null argument"; return CFD2D_EINVAL; }
    *out = nullptr;
    if (m->nc < 0 || m->nc_ex < m->nc || m->ne < 0 || p->nmat < 1 || p->nmat > 255 || p->nbc < 0) {
        g_create_error = "inconsistent sizes (nc, nc_ex, ne, nmat, nbc)";
        return CFD2D_EINVAL;
    }
    if (c->order != 1 && c->order != 2) { g_create_error = "ctrl.order must be 1 or 2"; return CFD2D_EINVAL; }
    if (c->flux != CFD2D_FLUX_GODUNOV && c->flux != CFD2D_FLUX_LAX) { g_create_error = "ctrl.flux must be GODUNOV or LAX"; return CFD2D_EINVAL; }
    const int nc = m->nc, nc_ex = m->nc_ex, ne = m->ne;
    // ---- validate topology on the host (cheap, once)
    for (int e = 0; e < ne; e++) {
        int c1 = m->edge_c1[e], c2 = m->edge_c2[e];
        if (c1 < 0 || c1 >= nc_ex || c2 < -1 || c2 >= nc_ex) { g_create_error = "edge references a cell out of range"; return CFD2D_EINVAL; }
        if (c2 < 0) {
            int ib = m->edge_bc[e];
            if (ib < 0 || ib >= p->nbc) {
                char b[128];
                snprintf(b, sizeof b, "Not defined boundary condition for edge %d", e);   // fvm_tvd.cpp:708
                g_create_error = b;
                return CFD2D_EBC;
            }
            int kind = p->bc_kind[ib];
            if (kind < CFD2D_BC_INLET || kind > CFD2D_BC_WALL) { g_create_error = "unknown boundary kind"; return CFD2D_EBC; }
        }
    }
    for (int i = 0; i < nc_ex; i++)
        if (m->cell_mat[i] < 0 || m->cell_mat[i] >= p->nmat) { g_create_error = "cell material index out of range"; return CFD2D_EINVAL; }
    // The per-cell slot order IS the summation order of the residual and of the gradients; the reference's
    // edge-ordered scatter (fvm_tvd.cpp:253-292, :353-363) corresponds to ascending edge ids, which its readers
    // produce (MeshReaderSalomeUnv.cpp:176-177,193-194).  A caller with another order would silently get
    // different bits, so the documented contract is enforced.
    for (int i = 0; i < nc; i++) {
        const int32_t* ce = m->cell_edges + 3 * (size_t)i;
        if (!(ce[0] < ce[1] && ce[1] < ce[2])) {
            char b[160];
            snprintf(b, sizeof b, "cell_edges of cell %d is not in ascending edge id (%d, %d, %d): the slot order is the summation order", i, ce[0], ce[1], ce[2]);
            g_create_error = b;
            return CFD2D_EINVAL;
        }
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        g_create_error = "no usable CUDA device (this path has no CPU fallback)";
        return CFD2D_ENODEV;
    }
    CUDA_TRY((cfd2d_fvm*)nullptr, cudaSetDevice(device));

    cfd2d_fvm* h = new cfd2d_fvm();
    h->device = device; h->nc = nc; h->nc_ex = nc_ex; h->ne = ne; h->nmat = p->nmat; h->nbc = p->nbc;
    h->ctrl = *c;
    if (h->ctrl.max_newton <= 0) h->ctrl.max_newton = 1000;
    h->TAU = c->TAU;
    int rc = 0;
#define TRY(x) do { rc = (x); if (rc) { g_create_error = h->error.empty() ? g_create_error : h->error; cfd2d_fvm_destroy(h); return rc; } } while (0)
    {
        cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete h; return CFD2D_ECUDA; }
        cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
    }
    // ---- device cell numbering: owned cells along a Hilbert curve (fvm_tiling.h); everything below
    // is built from the renumbered mesh `pm`; the caller's numbering only reappears in set/get_state,
    // the parity hooks and the sweep order of remediateLimCells.  Edge ids stay the caller's.
    bool hilbert = true;
    if (const char* ev = getenv("CFD2D_HILBERT")) hilbert = atoi(ev) != 0;
    hilbert_cell_order(nc, nc_ex, m->cell_cx, m->cell_cy, hilbert, h->perm, h->orig);
    HostMesh pm;
    permute_mesh(m, h->perm, h->orig, pm);
    // ---- internal edge order.  The flux kernel may visit edges in any order (each edge is independent,
    // F is addressed through the per-cell slot table, whose SLOT order -- the summation order -- is
    // untouched).  Edges are grouped by the direction of their normal (16 bins over [0, pi)), inner
    // edges first, boundary edges last, ascending id within a group.  Why: whether the two waves of
    // the Riemann problem are shocks or rarefactions is decided by the sign of (uL-uR).n, i.e. by the
    // local velocity gradient contracted twice with n -- spatially smooth for a fixed direction but
    // alternating between the three edge directions of a triangle.  Grouping by direction makes the
    // exp/log (rarefaction) vs sqrt/div (shock) branches of rim_orig warp-coherent; it also makes
    // the c1/c2 gathers of consecutive lanes monotone in memory.
    // The grouping is done inside tiles of EDGE_TILE consecutive edges, so that the cells a tile
    // touches (a few MB of W/G records) stay in L2 while its direction groups are swept one after
    // the other -- without tiling every group pass would stream all cell records from HBM again.
    const int NBIN = 16;
    int EDGE_TILE = 65536;
    if (const char* ev = getenv("CFD2D_EDGE_TILE")) EDGE_TILE = atoi(ev);   // 0 = keep the caller's order
    std::vector<int> order(ne), epos(ne);
    {
        const double PI_ = 3.14159265358979323846;
        std::vector<int> key(ne);
        for (int e = 0; e < ne; e++) {
            int b;
            if (pm.edge_c2[e] < 0) b = NBIN;                      // boundary edges: own group, last
            else {
                double a = atan2(pm.edge_ny[e], pm.edge_nx[e]);   // (-pi, pi]
                if (a < 0) a += PI_;                              // fold n and -n together
                b = (int)floor((a + PI_ / (2 * NBIN)) / (PI_ / NBIN));
                if (b >= NBIN || b < 0) b = 0;
            }
            key[e] = b;
        }
        // multi-rank: edges that touch a halo cell go last (their flux waits for the halo exchange)
        std::vector<int> ids;
        ids.reserve(ne);
        for (int e = 0; e < ne; e++) if (!(pm.edge_c1[e] >= nc || pm.edge_c2[e] >= nc)) ids.push_back(e);
        h->ne_int = (int)ids.size();
        for (int e = 0; e < ne; e++) if (pm.edge_c1[e] >= nc || pm.edge_c2[e] >= nc) ids.push_back(e);
        // edges are visited in the order of their c1 cell on the device (Hilbert order) instead of the
        // caller's edge order: consecutive edges gather neighbouring cell records and the staged fluxes
        // are laid out like the cells that gather them (B200, 4 M cells: k_flux 0.584 -> 0.559 ms,
        // k_update 0.133 -> 0.120 ms; profiles/r02d_edge_order.jsonl).  CFD2D_EDGE_SORT=0: caller's order
        int edge_sort = 1;
        if (const char* ev = getenv("CFD2D_EDGE_SORT")) edge_sort = atoi(ev);
        if (edge_sort) {
            auto key_of = [&](int e) {
                int k1 = pm.edge_c1[e], k2 = pm.edge_c2[e];
                return (edge_sort == 2 && k2 >= 0 && k2 < k1) ? k2 : k1;       // 1: c1, 2: the lower device id
            };
            auto by_cell = [&](int a, int b) {
                int ka = key_of(a), kb = key_of(b);
                return ka != kb ? ka < kb : a < b;
            };
            std::sort(ids.begin(), ids.begin() + h->ne_int, by_cell);
            std::sort(ids.begin() + h->ne_int, ids.end(), by_cell);
        }
        if (EDGE_TILE <= 0) { for (int q = 0; q < ne; q++) { order[q] = ids[q]; epos[ids[q]] = q; } }
        else {
            const int seg_beg[2] = {0, h->ne_int}, seg_end[2] = {h->ne_int, ne};
            for (int sg = 0; sg < 2; sg++)
                for (int t0 = seg_beg[sg]; t0 < seg_end[sg]; t0 += EDGE_TILE) {
                    int t1 = t0 + EDGE_TILE < seg_end[sg] ? t0 + EDGE_TILE : seg_end[sg];
                    int cnt[NBIN + 2] = {0};
                    for (int i = t0; i < t1; i++) cnt[key[ids[i]] + 1]++;
                    for (int b = 0; b <= NBIN; b++) cnt[b + 1] += cnt[b];
                    for (int i = t0; i < t1; i++) { int e = ids[i]; int q = t0 + cnt[key[e]]++; order[q] = e; epos[e] = q; }
                }
        }
    }
    h->edge_pos = epos;
    // ---- per (cell, slot) gather tables
    std::vector<int> s_nb(3 * (size_t)nc), s_es(3 * (size_t)nc);
    std::vector<double> s_nx(3 * (size_t)nc), s_ny(3 * (size_t)nc), s_l(3 * (size_t)nc);
    for (int cc = 0; cc < nc; cc++) {
        for (int k = 0; k < 3; k++) {
            int e = pm.cell_edges[3 * (size_t)cc + k];
            if (e < 0 || e >= ne) { g_create_error = "cell_edges entry out of range"; cfd2d_fvm_destroy(h); return CFD2D_EINVAL; }
            size_t o = (size_t)k * nc + cc;
            if (pm.edge_c1[e] == cc) {
                s_nb[o] = pm.edge_c2[e] >= 0 ? pm.edge_c2[e] : -1 - pm.edge_bc[e];
                s_nx[o] = pm.edge_nx[e]; s_ny[o] = pm.edge_ny[e];
                s_es[o] = epos[e] * 2;
            } else if (pm.edge_c2[e] == cc) {
                s_nb[o] = pm.edge_c1[e];
                s_nx[o] = -pm.edge_nx[e]; s_ny[o] = -pm.edge_ny[e];
                s_es[o] = epos[e] * 2 + 1;
            } else {
                g_create_error = "cell_edges names an edge that does not touch the cell";
                cfd2d_fvm_destroy(h);
                return CFD2D_EINVAL;
            }
            s_l[o] = pm.edge_l[e];
        }
    }
    // ---- per edge tables
    std::vector<int2> e_c(ne);
    std::vector<double2> e_n(ne);
    std::vector<double> e_l2(ne);
    std::vector<double4> e_d1(ne), e_d2(ne);
    std::vector<int> ebc(ne);
    for (int q = 0; q < ne; q++) {
        const int e = order[q];                                       // caller's edge id
        int c1 = pm.edge_c1[e], c2 = pm.edge_c2[e];
        e_c[q] = make_int2(c1, c2);
        e_n[q] = make_double2(pm.edge_nx[e], pm.edge_ny[e]);
        e_l2[q] = pm.edge_l[e] * 0.5;                                 // fvm_tvd.cpp:335
        ebc[q] = pm.edge_bc[e];
        const double* g = pm.edge_gp.data() + 4 * (size_t)e;
        // DL = PE - P(cell) (fvm_tvd.cpp:661-664): the same subtraction, done once
        e_d1[q] = make_double4(g[0] - pm.cell_cx[c1], g[1] - pm.cell_cy[c1], g[2] - pm.cell_cx[c1], g[3] - pm.cell_cy[c1]);
        if (c2 >= 0) e_d2[q] = make_double4(g[0] - pm.cell_cx[c2], g[1] - pm.cell_cy[c2], g[2] - pm.cell_cx[c2], g[3] - pm.cell_cy[c2]);
        else e_d2[q] = make_double4(0, 0, 0, 0);
    }
    std::vector<MatC> mats(p->nmat);
    for (int i = 0; i < p->nmat; i++) {
        MatC q;
        q.M = p->mat_M[i];
        q.Cv = p->mat_Cp[i] - CFD2D_GR / p->mat_M[i];   // global.cpp:11
        q.gam = p->mat_Cp[i] / q.Cv;                    // global.cpp:12
        q.gm1 = q.gam - 1;
        mats[i] = q;
    }
    std::vector<unsigned char> cmat(nc_ex);
    for (int i = 0; i < nc_ex; i++) cmat[i] = (unsigned char)pm.cell_mat[i];
    std::vector<int> bkind(p->bc_kind, p->bc_kind + p->nbc);
    std::vector<double> bpar(p->bc_par, p->bc_par + 4 * (size_t)p->nbc);
    const std::vector<double>& cS = pm.cell_S;

    KParams& P = h->P;
    P.nc = nc; P.nc_ex = nc_ex; P.ne = ne; P.nmat = p->nmat;
    P.order = c->order; P.flux = c->flux; P.max_newton = h->ctrl.max_newton; P.steady = c->steady;
    P.CFL = c->CFL;
    memcpy(P.lim, p->limits, sizeof P.lim);
    P.rim = make_rim(1.4);                               // double __GAM = 1.4; fvm_tvd.cpp:345
    TRY(dev_upload(h, &P.mat, mats));
    TRY(dev_upload(h, &P.bc_kind, bkind));
    TRY(dev_upload(h, &P.bc_par, bpar));
    TRY(dev_upload(h, &P.cell_mat, cmat));
    TRY(dev_upload(h, &P.s_nb, s_nb));
    TRY(dev_upload(h, &P.s_nx, s_nx));
    TRY(dev_upload(h, &P.s_ny, s_ny));
    TRY(dev_upload(h, &P.s_l, s_l));
    TRY(dev_upload(h, &P.s_es, s_es));
    TRY(dev_upload(h, &P.cell_S, cS));
    TRY(dev_upload(h, &P.e_c, e_c));
    TRY(dev_upload(h, &P.e_n, e_n));
    TRY(dev_upload(h, &P.e_l2, e_l2));
    TRY(dev_upload(h, &P.e_d1, e_d1));
    TRY(dev_upload(h, &P.e_d2, e_d2));
    TRY(dev_upload(h, &P.e_bc, ebc));
    TRY(dev_alloc(h, &P.cfl, (size_t)nc + 2));      // + 16 bytes: the pipelined kernel's bulk copies are whole 16-byte units
    TRY(dev_alloc(h, &P.ctau, (size_t)nc));
    TRY(dev_alloc(h, &P.flag, (size_t)nc + 4));
    TRY(dev_alloc(h, &h->err, 4));
    P.err = h->err;
    int cap = 1;
    while (cap < nc) cap <<= 1;
    P.lim_cap = cap;
    TRY(dev_alloc(h, &P.lim_list, (size_t)cap));
    TRY(dev_alloc(h, &h->Ua, (size_t)nc_ex));
    TRY(dev_alloc(h, &h->Ub, (size_t)nc_ex));
    TRY(dev_alloc(h, &h->W, (size_t)nc_ex));
    TRY(dev_alloc(h, &h->Wb, (size_t)nc_ex));
    TRY(dev_alloc(h, &h->io_u32, (size_t)nc));
    TRY(dev_upload(h, &P.c_perm, h->perm));
    TRY(dev_upload(h, &P.c_orig, h->orig));
    TRY(dev_alloc(h, &h->G, 2 * (size_t)(nc_ex ? nc_ex : 1)));
    TRY(dev_alloc(h, &h->F, (size_t)ne));
    for (int i = 0; i < 6; i++) TRY(dev_alloc(h, &h->io[i], (size_t)nc));
    TRY(dev_alloc(h, &h->tau_bits, 1));
    TRY(dev_alloc(h, &P.rstat, (size_t)nc_ex));
    {
        const size_t n1 = nc ? nc : 1, nx = nc_ex ? nc_ex : 1;
        TRY(cuda_rc(h, cudaMemset(h->err, 0, 4 * sizeof(int)), "cudaMemset(err)"));
        TRY(cuda_rc(h, cudaMemset(P.flag, 0, n1 * sizeof(unsigned int)), "cudaMemset(flag)"));
        TRY(cuda_rc(h, cudaMemset(P.rstat, 0, nx), "cudaMemset(rstat)"));
        TRY(cuda_rc(h, cudaMemset(h->Ua, 0, nx * sizeof(double4)), "cudaMemset(Ua)"));
        TRY(cuda_rc(h, cudaMemset(h->Ub, 0, nx * sizeof(double4)), "cudaMemset(Ub)"));
        TRY(cuda_rc(h, cudaMemset(h->W, 0, nx * sizeof(double4)), "cudaMemset(W)"));
        TRY(cuda_rc(h, cudaMemset(h->Wb, 0, nx * sizeof(double4)), "cudaMemset(Wb)"));
        TRY(cuda_rc(h, cudaMemset(h->G, 0, 2 * nx * sizeof(double4)), "cudaMemset(G)"));
        TRY(cuda_rc(h, cudaMemset(P.cfl, 0, n1 * sizeof(double)), "cudaMemset(cfl)"));
        TRY(cuda_rc(h, cudaMemset(P.ctau, 0, n1 * sizeof(double)), "cudaMemset(ctau)"));
    }
    // ---- step layout.  Default: three sweeps per stage.  Measured on B200 at 4 M cells
    // (profiles/README.md) the tile-fused kernel moves ~35 % fewer HBM bytes but its barrier-separated
    // phases expose more load latency than the three full-width sweeps hide; it stays selectable
    // (cfd2d_fvm_use_fused, CFD2D_FUSED=1) and is held to bit-identity with the sweeps by the tests.
    // Its plan (1-2 GB of tables at 4 M cells) is only built when it is selected.
    // ---- default step layout, from the 4 M-cell measurements in profiles/README.md: three sweeps for
    // every scheme but first-order Lax-Friedrichs (the single cell-parallel sweep k_cell_lf1, lf1_cell
    // below).  With the lane-split k_flux and 256-bit record accesses the sweeps run LF order 2 in
    // 0.99 ms per step against 1.11-1.18 for the pipelined tile kernel, which moves 2.5x fewer bytes but
    // is bound by its instruction stream; Godunov was always faster as three sweeps.
    h->fused = false;
    int layout = 0;
    bool layout_by_default = true;
    if (const char* ev = getenv("CFD2D_FUSED")) { layout = atoi(ev); layout_by_default = false; }
    h->fused = layout == 1;
    h->pm.reset(new HostMesh(std::move(pm)));
    const HostMesh& pmr = *h->pm;
    if (h->fused) TRY(build_fused_plan(h));
    if (const char* ev = getenv("CFD2D_EXACT_RIEMANN")) h->exact_riemann = atoi(ev) != 0;
    {
        int smc = 0;
        if (cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, h->device) == cudaSuccess) h->sm_count = smc;
    }
    h->lf1_cell = (c->flux == CFD2D_FLUX_LAX && c->order == 1);
    if (const char* ev = getenv("CFD2D_LF1_CELL")) h->lf1_cell = h->lf1_cell && atoi(ev) != 0;
    if (!(halo && halo->nranks > 1)) {
        if (const char* ev = getenv("CFD2D_DIAG_SPLIT")) h->diag_split = atoi(ev) != 0;
        if (h->diag_split) {
            std::vector<int> ci(nc);
            for (int i = 0; i < nc; i++) ci[i] = i;
            h->n_cells_int = nc; h->n_cells_bnd = 0;
            { const int* q = nullptr; TRY(dev_upload(h, &q, ci)); h->d_cells_int = (int*)q; }
        }
    }
    if (halo && halo->nranks > 1) {
        std::string herr;
        // the send lists name the caller's cells: translate to device ids
        int nsend = 0;
        for (int r = 0; r < halo->nranks; r++) nsend += halo->send_count[r];
        std::vector<int> send_dev(nsend > 0 ? nsend : 1, 0);
        for (int i = 0; i < nsend; i++) {
            int s = halo->send_ind[i];
            if (s < 0 || s >= nc) { g_create_error = "send_ind entry is not an owned cell"; cfd2d_fvm_destroy(h); return CFD2D_EINVAL; }
            send_dev[i] = h->perm[s];
        }
        // remediateLimCells across ranks: a flagged owned cell reads the PRE-remediation value of a halo
        // c2-neighbour, which equals the serial ascending sweep iff that neighbour has the higher global
        // id.  Decomp keeps the global edge orientation and the reference's readers create every edge from
        // its lower-numbered cell, so c1 < c2 holds there; verify it when the global ids are given.
        if (halo->cell_gid) {
            for (int e = 0; e < ne; e++) {
                const int a1 = m->edge_c1[e], a2 = m->edge_c2[e];
                if (a2 >= 0 && (a1 >= nc || a2 >= nc) && !(halo->cell_gid[a1] < halo->cell_gid[a2])) {
                    g_create_error = "multi-rank handle: an edge across the partition has global id(c1) > id(c2); "
                                     "remediateLimCells would not equal the serial sweep (fvm_tvd.cpp:464-499)";
                    cfd2d_fvm_destroy(h);
                    return CFD2D_EINVAL;
                }
            }
        }
        cfd2d_halo hd = *halo;
        hd.send_ind = send_dev.data();
        h->halo = halo_create(&hd, nc, nc_ex, device, &herr);
        if (!h->halo) { g_create_error = herr; cfd2d_fvm_destroy(h); return CFD2D_ENCCL; }
        {   // halo transport: direct peer stores over NVLink (CUDA IPC) when CFD2D_HALO_P2P=1 and every rank can
            // map its neighbours, else ncclSend/ncclRecv.  Collective: the variable must agree on all ranks.
            int want = 0;
            if (const char* ev = getenv("CFD2D_HALO_P2P")) want = atoi(ev);
            if (want) {
                double4* fields[3] = {h->Ua, h->Ub, h->G};
                int prc = halo_p2p_enable(h->halo, fields, 3, h->stream);
                if (prc) { g_create_error = halo_error(h->halo); cfd2d_fvm_destroy(h); return prc; }
            }
        }
        // Multi-rank steps are launched eagerly: the host thread stays well ahead of the 1.2 ms of device
        // work per step, and the captured two-stream graph (fork/join through events, NCCL point-to-point
        // nodes) replays 3-4 % SLOWER than the same launches made directly (2 and 4 B200: 1.280 vs 1.243,
        // 1.296 vs 1.243 ms per step, profiles/README.md section 4).  CFD2D_GRAPH_MULTI=1 captures it.
        h->use_graph = false;
        if (const char* ev = getenv("CFD2D_GRAPH_MULTI")) h->use_graph = atoi(ev) != 0;
        h->n_send = nsend;
        { const int* q = nullptr; TRY(dev_upload(h, &q, send_dev)); h->d_send_dev = (int*)q; }
        // owned cells without / with a halo neighbour (gradient sweep split, enqueue_step_unfused)
        std::vector<int> ci, cb;
        for (int cc = 0; cc < nc; cc++) {
            bool b = false;
            for (int k = 0; k < 3; k++) {
                int e = pmr.cell_edges[3 * (size_t)cc + k];
                int nb = pmr.edge_c1[e] == cc ? pmr.edge_c2[e] : pmr.edge_c1[e];
                if (nb >= nc) b = true;
            }
            (b ? cb : ci).push_back(cc);
        }
        h->n_cells_int = (int)ci.size(); h->n_cells_bnd = (int)cb.size();
        { const int* q = nullptr; TRY(dev_upload(h, &q, ci)); h->d_cells_int = (int*)q; }
        { const int* q = nullptr; TRY(dev_upload(h, &q, cb)); h->d_cells_bnd = (int*)q; }
        if (const char* ev = getenv("CFD2D_OVERLAP")) h->overlap = atoi(ev) != 0;
        if (const char* ev = getenv("CFD2D_DIAG_NO_EXCHANGE")) h->skip_exchange = atoi(ev) != 0;
        if (const char* ev = getenv("CFD2D_COLLECTIVE_ERRORS")) h->collective_errors = atoi(ev) != 0;
        {   // highest priority: the small pack / NCCL kernels must not queue behind a full-GPU sweep
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            TRY(cuda_rc(h, cudaStreamCreateWithPriority(&h->comm, cudaStreamNonBlocking, hi), "cudaStreamCreateWithPriority"));
        }
        TRY(cuda_rc(h, cudaEventCreateWithFlags(&h->ev_G, cudaEventDisableTiming), "cudaEventCreateWithFlags"));
        TRY(cuda_rc(h, cudaEventCreateWithFlags(&h->ev_stage, cudaEventDisableTiming), "cudaEventCreateWithFlags"));
        TRY(cuda_rc(h, cudaEventCreateWithFlags(&h->ev_U, cudaEventDisableTiming), "cudaEventCreateWithFlags"));
    }
    if (layout == 2) {
        int prc = build_pipe_plan_dev(h);
        if (prc == CFD2D_OK) h->pipe = true;
        else if (layout_by_default && prc == CFD2D_EINVAL) h->error.clear();   // e.g. a tile of an unordered mesh (CFD2D_HILBERT=0)
                                                                               // whose rings exceed shared memory: three sweeps
        else TRY(prc);
    }
    TRY(cuda_rc(h, cudaDeviceSynchronize(), "cudaDeviceSynchronize (create)"));
#undef TRY
    *out = h;
    return CFD2D_OK;
}

void cfd2d_fvm_destroy(cfd2d_fvm* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_graph(h);
    if (h->halo) halo_destroy(h->halo);
    for (void* p : h->allocs) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_G) cudaEventDestroy(h->ev_G);
    if (h->ev_stage) cudaEventDestroy(h->ev_stage);
    if (h->ev_U) cudaEventDestroy(h->ev_U);
    if (h->comm) { cudaStreamSynchronize(h->comm); cudaStreamDestroy(h->comm); }
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->ev_snap) cudaEventDestroy(h->ev_snap);
    if (h->ev_snap_done) cudaEventDestroy(h->ev_snap_done);
    if (h->snap_host) cudaFreeHost(h->snap_host);
    if (h->gather_buf) cudaFree(h->gather_buf);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

int cfd2d_fvm_set_stream(cfd2d_fvm* h, void* s) {
    if (!h) return CFD2D_EINVAL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    drop_graph(h);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)s;
    h->own_stream = false;
    return 0;
}

int cfd2d_fvm_use_graph(cfd2d_fvm* h, int on) {
    if (!h) return CFD2D_EINVAL;
    h->use_graph = on != 0;
    if (!h->use_graph) drop_graph(h);
    return 0;
}

int cfd2d_fvm_use_fused(cfd2d_fvm* h, int on) {
    if (!h) return CFD2D_EINVAL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->comm) cudaStreamSynchronize(h->comm);
    drop_graph(h);
    if (on != 0 && on != 1 && on != 2) { h->error = "layout must be 0 (three sweeps), 1 (k_stage) or 2 (k_stage_pipe)"; return CFD2D_EINVAL; }
    if (on == 1) { int rc = build_fused_plan(h); if (rc) return rc; }
    if (on == 2) { int rc = build_pipe_plan_dev(h); if (rc) return rc; }
    if (on != 2) ensure_W(h);                 // back to a layout that reads the primitive cache
    h->fused = on == 1;
    h->pipe = on == 2;
    return 0;
}

int cfd2d_fvm_use_exact_riemann(cfd2d_fvm* h, int on) {
    if (!h) return CFD2D_EINVAL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->comm) cudaStreamSynchronize(h->comm);
    drop_graph(h);
    h->exact_riemann = on != 0;
    if (h->have_plan)
        for (int st = 1; st <= 2; st++)
            CUDA_TRY(h, cudaFuncSetAttribute((const void*)stage_kernel(h, st), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->stage_smem));
    return 0;
}

const char* cfd2d_fvm_plan_summary(const cfd2d_fvm* h) { return h ? h->plan_summary.c_str() : ""; }

int cfd2d_tiling_plan(const cfd2d_mesh* m, int tile_cells, int hilbert, int32_t* perm_out, int64_t* stats_out) {
    g_create_error.clear();
    if (!m || m->nc < 0 || m->nc_ex < m->nc || m->ne < 0) { g_create_error = "bad mesh"; return CFD2D_EINVAL; }
    std::vector<int> perm, orig;
    hilbert_cell_order(m->nc, m->nc_ex, m->cell_cx, m->cell_cy, hilbert != 0, perm, orig);
    HostMesh pm;
    permute_mesh(m, perm, orig, pm);
    TilePlan tp;
    std::string err = build_tile_plan(pm, tile_cells, tp);
    if (!err.empty()) { g_create_error = err; return CFD2D_EINVAL; }
    // ---- invariants the kernel relies on
    for (int t = 0; t < tp.ntiles && err.empty(); t++) {
        const TileInfo& ti = tp.tiles[t];
        auto gid = [&](int l) { return l < ti.n_own ? ti.cbeg + l : tp.ring[ti.roff + l - ti.n_own]; };
        for (int q = 0; q < ti.ne_t; q++) {
            size_t eo = (size_t)ti.eoff + q;
            int e = tp.e_id[eo];
            int l1 = (int)(tp.e_cl[eo] & 0xffffu), l2 = (int)(tp.e_cl[eo] >> 16);
            if (l1 >= ti.n_l || gid(l1) != pm.edge_c1[e] || tp.e_c1[eo] != pm.edge_c1[e]) { err = "edge c1 local/global id mismatch"; break; }
            if (pm.edge_c2[e] >= 0) {
                if (l2 >= ti.n_l || gid(l2) != pm.edge_c2[e] || tp.e_c2[eo] != pm.edge_c2[e]) { err = "edge c2 local/global id mismatch"; break; }
            } else if (tp.e_c2[eo] != -1 - pm.edge_bc[e] || l2 != 0xffff || l1 >= ti.n_own) { err = "boundary edge encoding mismatch"; break; }
        }
        for (int j = 0; j < ti.n_own && err.empty(); j++) {
            int c = ti.cbeg + j;
            for (int k = 0; k < 3; k++) {
                int es = tp.u_es[(size_t)k * pm.nc + c];
                int q = es >> 1;
                if (q < 0 || q >= ti.ne_t) { err = "u_es out of the tile's edge range"; break; }
                int e = tp.e_id[(size_t)ti.eoff + q];
                if (e != pm.cell_edges[3 * (size_t)c + k]) { err = "u_es names the wrong edge"; break; }
                if (((es & 1) ? pm.edge_c2[e] : pm.edge_c1[e]) != c) { err = "u_es side bit wrong"; break; }
            }
        }
        for (int j = 0; j < ti.n_g && err.empty(); j++) {
            int c = gid(j);
            if (c < 0 || c >= pm.nc) { err = "gradient computed for a non-owned cell"; break; }
            for (int k = 0; k < 3; k++) {
                int e = pm.cell_edges[3 * (size_t)c + k];
                int nb = tp.g_nb[(size_t)ti.goff + (size_t)k * ti.gstride + j];
                int want = pm.edge_c1[e] == c ? pm.edge_c2[e] : pm.edge_c1[e];
                if (want < 0) want = -1 - pm.edge_bc[e];
                if (nb != want) { err = "gradient table neighbour mismatch"; break; }
            }
        }
        for (int j = ti.n_g; j < ti.n_l && err.empty(); j++)
            if (gid(j) < pm.nc) err = "computable ring-1 cell listed as halo";
    }
    if (!err.empty()) { g_create_error = "tile plan invariant violated: " + err; return CFD2D_EINVAL; }
    if (perm_out) for (int i = 0; i < m->nc_ex; i++) perm_out[i] = perm[i];
    if (stats_out) {
        stats_out[0] = tp.ntiles; stats_out[1] = tp.nl_max; stats_out[2] = tp.ne_max; stats_out[3] = tp.sum_ng;
        stats_out[4] = tp.sum_ne; stats_out[5] = tp.sum_ring; stats_out[6] = (int64_t)tp.interior.size();
        stats_out[7] = (int64_t)tp.boundary.size();
    }
    return 0;
}

int cfd2d_pipe_plan(const cfd2d_mesh* m, int tile_cells, int dir_bins, int hilbert, int64_t* stats_out) {
    g_create_error.clear();
    if (!m || m->nc < 0 || m->nc_ex < m->nc || m->ne < 0) { g_create_error = "bad mesh"; return CFD2D_EINVAL; }
    std::vector<int> perm, orig;
    hilbert_cell_order(m->nc, m->nc_ex, m->cell_cx, m->cell_cy, hilbert != 0, perm, orig);
    HostMesh pm;
    permute_mesh(m, perm, orig, pm);
    PipePlan pp;
    std::string err = build_pipe_plan(pm, tile_cells, dir_bins != 0, pp);
    if (!err.empty()) { g_create_error = err; return CFD2D_EINVAL; }
    // ---- invariants k_stage_pipe relies on, re-derived from the blob bytes alone
    long long owned = 0;
    for (int t = 0; t < pp.ntiles && err.empty(); t++) {
        const PipeTile& ti = pp.tiles[t];
        const unsigned char* b = pp.blob.data() + ((size_t)ti.blob_off << 4);
        auto gid = [&](int l) { return l < ti.n_own ? ti.cbeg + l : pp.ring[ti.roff + l - ti.n_own]; };
        if (ti.cbeg != owned || (ti.blob_bytes & 15) || (ti.cbeg & 7)) { err = "tile range / alignment"; break; }
        owned += ti.n_own;
        const int S = (ti.n_own + 7) & ~7, R = ti.n_g - ti.n_own;
        const uint16_t* slot = reinterpret_cast<const uint16_t*>(b + ti.o_slot);
        const double* en = reinterpret_cast<const double*>(b + ti.o_en);
        const double* egp = reinterpret_cast<const double*>(b + ti.o_egp);
        const double* cxy = reinterpret_cast<const double*>(b + ti.o_cxy);
        const double* Sv = reinterpret_cast<const double*>(b + ti.o_S);
        for (int l = 0; l < ti.n_l && err.empty(); l++)
            if (cxy[2 * l] != pm.cell_cx[gid(l)] || cxy[2 * l + 1] != pm.cell_cy[gid(l)]) err = "cell centre";
        for (int l = 0; l < ti.n_g && err.empty(); l++) {
            if (Sv[l] != pm.cell_S[gid(l)]) err = "cell area";
            if (gid(l) >= pm.nc) err = "gradient computed for a non-owned cell";
        }
        for (int l = ti.n_g; l < ti.n_l && err.empty(); l++) if (gid(l) < pm.nc) err = "owned ring-1 cell listed as halo";
        for (int l = 0; l < ti.n_l2 && err.empty(); l++) if (b[ti.o_mat + l] != (unsigned char)pm.cell_mat[gid(l)]) err = "material";
        for (int j = 0; j < ti.n_own && err.empty(); j++) {
            const int c = ti.cbeg + j;
            for (int k = 0; k < 3; k++) {
                const int es = slot[k * S + j], q = es >> 1, e = pm.cell_edges[3 * (size_t)c + k];
                if (q >= ti.ne_t) { err = "slot out of the tile's edge range"; break; }
                double le; uint32_t cl;
                memcpy(&le, b + ti.o_el + 16 * (size_t)q, 8); memcpy(&cl, b + ti.o_el + 16 * (size_t)q + 8, 4);
                const int l1 = (int)(cl & 0xffffu), l2 = (int)(cl >> 16);
                if (le != pm.edge_l[e] || en[2 * q] != pm.edge_nx[e] || en[2 * q + 1] != pm.edge_ny[e]) { err = "slot names the wrong edge (geometry)"; break; }
                for (int i = 0; i < 4; i++) if (egp[4 * (size_t)q + i] != pm.edge_gp[4 * (size_t)e + i]) err = "Gauss points";
                if (gid(l1) != pm.edge_c1[e]) { err = "edge c1 local id"; break; }
                if (pm.edge_c2[e] >= 0 ? (l2 >= ti.n_l || gid(l2) != pm.edge_c2[e]) : (l2 != (0xff00 | pm.edge_bc[e]))) { err = "edge c2 local id / bc"; break; }
                if (((es & 1) ? pm.edge_c2[e] : pm.edge_c1[e]) != c) { err = "slot side bit"; break; }
            }
        }
        const int* gnb = reinterpret_cast<const int*>(b + ti.o_gnb);
        const double* gn = reinterpret_cast<const double*>(b + ti.o_gn);
        for (int r = 0; r < R && err.empty(); r++) {
            const int c = gid(ti.n_own + r);
            for (int k = 0; k < 3; k++) {
                const int e = pm.cell_edges[3 * (size_t)c + k];
                const bool is1 = pm.edge_c1[e] == c;
                const int want = is1 ? pm.edge_c2[e] : pm.edge_c1[e];
                const int nb = gnb[k * R + r];
                if (want < 0 ? nb != -1 - pm.edge_bc[e] : (nb < 0 || nb >= ti.n_l2 || gid(nb) != want)) { err = "ring-1 gradient neighbour"; break; }
                if (gn[(k * 3 + 0) * R + r] != (is1 ? pm.edge_nx[e] : -pm.edge_nx[e]) || gn[(k * 3 + 2) * R + r] != pm.edge_l[e]) { err = "ring-1 gradient geometry"; break; }
            }
        }
    }
    if (err.empty() && owned != pm.nc) err = "tiles do not cover the owned cells";
    if (!err.empty()) { g_create_error = "pipe plan invariant violated: " + err; return CFD2D_EINVAL; }
    if (stats_out) {
        stats_out[0] = pp.ntiles; stats_out[1] = pp.nl2_max; stats_out[2] = pp.ne_max; stats_out[3] = pp.blob_max;
        stats_out[4] = pp.sum_ne; stats_out[5] = pp.sum_ring1; stats_out[6] = pp.sum_ring2; stats_out[7] = (int64_t)pp.blob.size();
        stats_out[8] = (int64_t)pp.interior.size(); stats_out[9] = (int64_t)pp.boundary.size();
        stats_out[10] = pp.nring_max; stats_out[11] = pp.nl_max;
    }
    return 0;
}

int cfd2d_fvm_set_state(cfd2d_fvm* h, const double* ro, const double* ru, const double* rv, const double* re,
                        const uint32_t* flag) {
    if (!h || !ro || !ru || !rv || !re) return CFD2D_EINVAL;
    NvtxScope nv("cfd2d_fvm_set_state (H2D + pack)");
    CUDA_TRY(h, cudaSetDevice(h->device));
    size_t n = (size_t)h->nc * sizeof(double);
    const double* src[4] = {ro, ru, rv, re};
    for (int i = 0; i < 4; i++) CUDA_TRY(h, cudaMemcpyAsync(h->io[i], src[i], n, cudaMemcpyHostToDevice, h->stream));
    if (flag && h->nc) {
        CUDA_TRY(h, cudaMemcpyAsync(h->io_u32, flag, (size_t)h->nc * 4, cudaMemcpyHostToDevice, h->stream));
        h->launches++;
        k_scatter_perm<uint32_t><<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->io_u32, h->P.flag);
    } else CUDA_TRY(h, cudaMemsetAsync(h->P.flag, 0, (size_t)(h->nc ? h->nc : 1) * 4, h->stream));
    if (h->nc) {
        h->launches++;
        k_pack_state<<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->io[0], h->io[1], h->io[2], h->io[3], h->Ua);
    }
    launch_prim(h, h->Ua, h->W, 0, h->nc, h->stream);
    if (h->halo) {                                     // all NCCL traffic of a handle goes through its comm stream
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        int rc = exchange_U(h, h->Ua, h->W, h->comm);
        if (rc) return rc;
        CUDA_TRY(h, cudaStreamSynchronize(h->comm));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int cfd2d_fvm_calc_time_step(cfd2d_fvm* h, double* tau_out) {
    if (!h) return CFD2D_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    ensure_W(h);
    if (h->ctrl.steady) {
        launch_tau_steady(h);
    } else {
        unsigned long long bits;
        memcpy(&bits, &h->TAU, 8);
        CUDA_TRY(h, cudaMemcpyAsync(h->tau_bits, &bits, 8, cudaMemcpyHostToDevice, h->stream));
        if (h->nc) {
            h->launches++;
            k_tau_min<<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->P, h->W, h->tau_bits);
        }
        CUDA_TRY(h, cudaMemcpyAsync(&bits, h->tau_bits, 8, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        double tau;
        memcpy(&tau, &bits, 8);
        if (h->halo) {
            int rc = halo_allreduce_min(h->halo, &tau, h->stream);   // MPI_Allreduce(MIN) analogue (fem_rkdg.cpp:413)
            if (rc) { h->error = halo_error(h->halo); return rc; }
        }
        h->TAU = tau;
        if (h->nc) {
            h->launches++;
            k_tau_fill<<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->P, tau);
        }
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    if (tau_out) *tau_out = h->TAU;
    return 0;
}

int cfd2d_fvm_step_async(cfd2d_fvm* h, int nsteps) {
    if (!h || nsteps < 0) return CFD2D_EINVAL;
    NvtxScope nv("cfd2d_fvm_step (RK2 steps: FVM_TVD::run loop body)");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int done = 0;
    if (h->use_graph && !h->profiling && h->halo)
        for (; done < nsteps && h->eager_steps < 2; done++, h->eager_steps++) {   // NCCL connections, lazy allocations
            int rc = enqueue_step(h);
            if (rc) return rc;
        }
    if (h->use_graph && !h->profiling) {
        if (!h->graph_exec && done < nsteps) {
            int64_t l0 = h->launches;
            CUDA_TRY(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_step(h);
            cudaError_t e = cudaStreamEndCapture(h->stream, &h->graph);
            h->graph_launches = (int)(h->launches - l0);
            h->launches = l0;
            if (rc || e != cudaSuccess) { drop_graph(h); cudaGetLastError(); }
            if (rc) return rc;
            CUDA_TRY(h, e);
            CUDA_TRY(h, cudaGraphInstantiate(&h->graph_exec, h->graph, 0));
        }
        for (int s = done; s < nsteps; s++) {
            CUDA_TRY(h, cudaGraphLaunch(h->graph_exec, h->stream));
            h->launches += h->graph_launches;
        }
    } else {
        for (int s = done; s < nsteps; s++) {
            int rc = enqueue_step(h);
            if (rc) return rc;
        }
    }
    if (!h->ctrl.steady) for (int s = 0; s < nsteps; s++) h->t += h->TAU;   // t += TAU, fvm_tvd.cpp:313
    return 0;
}

int cfd2d_fvm_sync(cfd2d_fvm* h) {
    if (!h) return CFD2D_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->comm) CUDA_TRY(h, cudaStreamSynchronize(h->comm));
    CUDA_TRY(h, cudaGetLastError());
    return check_device_errors(h);
}

int cfd2d_fvm_step(cfd2d_fvm* h, int nsteps) {
    int rc = cfd2d_fvm_step_async(h, nsteps);
    if (rc) return rc;
    return cfd2d_fvm_sync(h);
}

int cfd2d_fvm_get_state(cfd2d_fvm* h, double* ro, double* ru, double* rv, double* re, double* cTau, uint32_t* flag) {
    if (!h || !ro || !ru || !rv || !re) return CFD2D_EINVAL;
    NvtxScope nv("cfd2d_fvm_get_state (unpack + D2H)");
    CUDA_TRY(h, cudaSetDevice(h->device));
    size_t n = (size_t)h->nc * sizeof(double);
    if (h->nc) {
        h->launches++;
        k_unpack_state<<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->Ua, h->io[0], h->io[1], h->io[2], h->io[3]);
    }
    double* dst[4] = {ro, ru, rv, re};
    for (int i = 0; i < 4; i++) CUDA_TRY(h, cudaMemcpyAsync(dst[i], h->io[i], n, cudaMemcpyDeviceToHost, h->stream));
    if (cTau && h->nc) {
        h->launches++;
        k_gather_perm<double><<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->P.ctau, h->io[4]);
        CUDA_TRY(h, cudaMemcpyAsync(cTau, h->io[4], n, cudaMemcpyDeviceToHost, h->stream));
    }
    if (flag && h->nc) {
        h->launches++;
        k_gather_perm<uint32_t><<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->P.flag, h->io_u32);
        CUDA_TRY(h, cudaMemcpyAsync(flag, h->io_u32, (size_t)h->nc * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

// ---- FVM_TVD::save without stalling the time loop ---------------------------
// begin: the state (caller order), cTau and the flags are unpacked on the compute stream into the
// handle's staging arrays -- ordered before whatever steps the caller enqueues next -- and copied to
// pinned host memory by a SEPARATE copy stream, so the D2H transfer and the caller's VTK writer run
// under the next chunk of steps.  end: waits for the copy and hands the arrays to the caller.
int cfd2d_fvm_snapshot_begin(cfd2d_fvm* h) {
    if (!h) return CFD2D_EINVAL;
    if (h->snap_pending) { h->error = "snapshot_begin: the previous snapshot was not collected (snapshot_end)"; return CFD2D_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t nc = (size_t)h->nc, nb = nc * sizeof(double);
    if (!h->copy_stream) {
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_snap_done, cudaEventDisableTiming));
        CUDA_TRY(h, cudaHostAlloc((void**)&h->snap_host, (5 * nb + nc * sizeof(uint32_t)) + 64, cudaHostAllocDefault));
    }
    if (h->nc) {
        h->launches += 3;
        k_unpack_state<<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->Ua, h->io[0], h->io[1], h->io[2], h->io[3]);
        k_gather_perm<double><<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->P.ctau, h->io[4]);
        k_gather_perm<uint32_t><<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->P.flag, h->io_u32);
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_snap, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_snap, 0));
    for (int i = 0; i < 5; i++)
        if (nb) CUDA_TRY(h, cudaMemcpyAsync(h->snap_host + (size_t)i * nc, h->io[i], nb, cudaMemcpyDeviceToHost, h->copy_stream));
    if (nc) CUDA_TRY(h, cudaMemcpyAsync(h->snap_host + 5 * nc, h->io_u32, nc * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->copy_stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_snap_done, h->copy_stream));
    // the staging arrays are reused by set_state / get_state / get_primitive: those wait for the copy
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_snap_done, 0));
    h->snap_pending = true;
    return 0;
}

int cfd2d_fvm_snapshot_end(cfd2d_fvm* h, double* ro, double* ru, double* rv, double* re, double* cTau, uint32_t* flag) {
    if (!h) return CFD2D_EINVAL;
    if (!h->snap_pending) { h->error = "snapshot_end without snapshot_begin"; return CFD2D_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaEventSynchronize(h->ev_snap_done));
    h->snap_pending = false;
    const size_t nc = (size_t)h->nc, nb = nc * sizeof(double);
    double* dst[5] = {ro, ru, rv, re, cTau};
    for (int i = 0; i < 5; i++) if (dst[i] && nb) memcpy(dst[i], h->snap_host + (size_t)i * nc, nb);
    if (flag && nc) memcpy(flag, h->snap_host + 5 * nc, nc * sizeof(uint32_t));
    return 0;
}

// Multi-rank: the owned-cell state of every rank collected on `root` (what a parallel Method does with
// Parallel::send/recv before the root writes the result file).  Rank-major concatenation on root.
int cfd2d_fvm_gather_state(cfd2d_fvm* h, int root, const int32_t* counts, double* ro, double* ru, double* rv, double* re,
                           double* cTau, uint32_t* flag) {
    if (!h || !counts) return CFD2D_EINVAL;
    if (!h->halo) { h->error = "gather_state needs a multi-rank handle"; return CFD2D_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int nr = halo_nranks(h->halo), me = halo_rank(h->halo);
    if (counts[me] != h->nc) { h->error = "gather_state: counts[rank] != owned cells of this handle"; return CFD2D_EINVAL; }
    const size_t REC = 5 * sizeof(double) + sizeof(uint32_t);          // bytes per cell: ro ru rv re cTau | flag
    std::vector<size_t> bytes(nr), off(nr + 1, 0);
    for (int p = 0; p < nr; p++) { bytes[p] = (size_t)counts[p] * REC; off[p + 1] = off[p] + bytes[p]; }
    const size_t need = (me == root ? off[nr] : bytes[me]) + 16;
    if (h->gather_cap < need) {
        if (h->gather_buf) cudaFree(h->gather_buf);
        h->gather_buf = nullptr; h->gather_cap = 0;
        CUDA_TRY(h, cudaMalloc((void**)&h->gather_buf, need));
        h->gather_cap = need;
    }
    char* mine = h->gather_buf + (me == root ? off[me] : 0);
    const size_t nc = (size_t)h->nc, nb = nc * sizeof(double);
    if (nc) {
        h->launches += 3;
        k_unpack_state<<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->Ua, h->io[0], h->io[1], h->io[2], h->io[3]);
        k_gather_perm<double><<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->P.ctau, h->io[4]);
        k_gather_perm<uint32_t><<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->P.flag, h->io_u32);
        for (int i = 0; i < 5; i++) CUDA_TRY(h, cudaMemcpyAsync(mine + (size_t)i * nb, h->io[i], nb, cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(mine + 5 * nb, h->io_u32, nc * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    int rc = halo_gather_bytes(h->halo, root, mine, bytes[me], h->gather_buf, bytes.data(), h->comm);   // NCCL traffic: comm stream
    if (rc) { h->error = halo_error(h->halo); return rc; }
    CUDA_TRY(h, cudaStreamSynchronize(h->comm));
    if (me == root) {
        std::vector<char> host(off[nr]);
        CUDA_TRY(h, cudaMemcpy(host.data(), h->gather_buf, off[nr], cudaMemcpyDeviceToHost));
        size_t c0 = 0;
        double* dst[5] = {ro, ru, rv, re, cTau};
        for (int p = 0; p < nr; p++) {
            const size_t n = (size_t)counts[p];
            const char* b = host.data() + off[p];
            for (int i = 0; i < 5; i++) if (dst[i] && n) memcpy(dst[i] + c0, b + (size_t)i * n * sizeof(double), n * sizeof(double));
            if (flag && n) memcpy(flag + c0, b + 5 * n * sizeof(double), n * sizeof(uint32_t));
            c0 += n;
        }
    }
    return 0;
}

int cfd2d_fvm_get_primitive(cfd2d_fvm* h, double* r, double* p, double* T, double* u, double* v, double* cz) {
    if (!h) return CFD2D_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    double* dst[6] = {r, p, T, u, v, cz};
    double* dev[6];
    for (int i = 0; i < 6; i++) dev[i] = dst[i] ? h->io[i] : nullptr;
    if (h->nc) {
        h->launches++;
        k_primitive_out<<<nblk(h->nc, 256), 256, 0, h->stream>>>(h->P, h->Ua, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5]);
    }
    for (int i = 0; i < 6; i++)
        if (dst[i]) CUDA_TRY(h, cudaMemcpyAsync(dst[i], h->io[i], (size_t)h->nc * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

double cfd2d_fvm_tau(const cfd2d_fvm* h) { return h ? h->TAU : 0.0; }
double cfd2d_fvm_time(const cfd2d_fvm* h) { return h ? h->t : 0.0; }
int64_t cfd2d_fvm_launch_count(const cfd2d_fvm* h) { return h ? h->launches : 0; }
int cfd2d_fvm_halo_transport(const cfd2d_fvm* h) { return (!h || !h->halo) ? 0 : (halo_p2p_active(h->halo) ? 2 : 1); }

int cfd2d_fvm_calc_grad(cfd2d_fvm* h, double* grad8) {
    if (!h || !grad8) return CFD2D_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    ensure_W(h);
    launch_grad(h);
    if (!h->grad_tmp) { int rc = dev_alloc(h, &h->grad_tmp, 8 * (size_t)(h->nc ? h->nc : 1)); if (rc) return rc; }
    if (h->nc) k_unpack_grad<<<nblk(2 * (long long)h->nc, 256), 256, 0, h->stream>>>(h->nc, h->P.c_perm, h->G, h->grad_tmp);
    CUDA_TRY(h, cudaMemcpyAsync(grad8, h->grad_tmp, (size_t)h->nc * 64, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int cfd2d_fvm_edge_fluxes(cfd2d_fvm* h, double* flux4) {
    if (!h || !flux4) return CFD2D_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    ensure_W(h);
    if (h->ctrl.order == 2) {
        launch_grad(h);
        if (h->halo) {                                 // all NCCL traffic of a handle goes through its comm stream
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            if ((rc = exchange_G(h, h->comm))) return rc;
            CUDA_TRY(h, cudaStreamSynchronize(h->comm));
        }
    }
    launch_flux(h, h->Ua, 0);
    std::vector<double> tmp(4 * (size_t)h->ne);
    CUDA_TRY(h, cudaMemcpyAsync(tmp.data(), h->F, (size_t)h->ne * 32, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    for (int e = 0; e < h->ne; e++) memcpy(flux4 + 4 * (size_t)e, tmp.data() + 4 * (size_t)h->edge_pos[e], 32);
    return check_device_errors(h);
}

int cfd2d_fvm_profile(cfd2d_fvm* h, int nsteps, double* ms, int64_t* launches) {
    if (!h || nsteps < 0) return CFD2D_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < CFD2D_NKERNELS; i++) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
    h->profiling = true;
    int rc = cfd2d_fvm_step_async(h, nsteps);
    h->profiling = false;
    if (rc) return rc;
    rc = cfd2d_fvm_sync(h);
    for (int i = 0; i < CFD2D_NKERNELS; i++) {
        if (ms) ms[i] = h->prof_ms[i];
        if (launches) launches[i] = h->prof_n[i];
    }
    return rc;
}

// ---- function-level known-answer entry points ---------------------------------------------------
static int kat_rim_impl(int device, int n, const double* in8, double gam, int max_newton, int fast, double* out5, int32_t* iters);

int cfd2d_kat_rim_orig(int device, int n, const double* in8, double gam, int max_newton, double* out5, int32_t* iters) {
    return kat_rim_impl(device, n, in8, gam, max_newton, 0, out5, iters);
}

int cfd2d_kat_rim_orig_fast(int device, int n, const double* in8, int max_newton, double* out5, int32_t* iters) {
    return kat_rim_impl(device, n, in8, 1.4, max_newton, 1, out5, iters);   // x^(1/7), x^(5/2): g = 1.4 only (fvm_tvd.cpp:345)
}

static int kat_rim_impl(int device, int n, const double* in8, double gam, int max_newton, int fast, double* out5, int32_t* iters) {
    g_create_error.clear();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); g_create_error = "no usable CUDA device"; return CFD2D_ENODEV; }
    if (n <= 0) return 0;
    if (max_newton <= 0) max_newton = 1000;
    cfd2d_fvm* none = nullptr;
    CUDA_TRY(none, cudaSetDevice(device));
    double *din = nullptr, *dout = nullptr; int* dit = nullptr;
    CUDA_TRY(none, cudaMalloc(&din, (size_t)n * 64));
    CUDA_TRY(none, cudaMalloc(&dout, (size_t)n * 40));
    CUDA_TRY(none, cudaMalloc(&dit, (size_t)n * 4));
    CUDA_TRY(none, cudaMemcpy(din, in8, (size_t)n * 64, cudaMemcpyHostToDevice));
    k_kat_rim<<<nblk(n, 128), 128>>>(make_rim(gam), max_newton, fast, n, din, dout, dit);
    CUDA_TRY(none, cudaDeviceSynchronize());
    CUDA_TRY(none, cudaMemcpy(out5, dout, (size_t)n * 40, cudaMemcpyDeviceToHost));
    std::vector<int> it(n);
    CUDA_TRY(none, cudaMemcpy(it.data(), dit, (size_t)n * 4, cudaMemcpyDeviceToHost));
    cudaFree(din); cudaFree(dout); cudaFree(dit);
    int bad = 0;
    for (int i = 0; i < n; i++) { if (iters) iters[i] = it[i]; if (it[i] < 0) bad = CFD2D_ENEWTON; }
    return bad;
}

int cfd2d_kat_urs(int device, int n, double M, double Cp, int mode, double* io8) {
    g_create_error.clear();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); g_create_error = "no usable CUDA device"; return CFD2D_ENODEV; }
    if (mode < 0 || mode > 2) { g_create_error = "URS mode must be 0, 1 or 2"; return CFD2D_EINVAL; }
    if (n <= 0) return 0;
    cfd2d_fvm* none = nullptr;
    CUDA_TRY(none, cudaSetDevice(device));
    MatC q;
    q.M = M;
    q.Cv = Cp - CFD2D_GR / M;     // global.cpp:11
    q.gam = Cp / q.Cv;            // global.cpp:12
    q.gm1 = q.gam - 1;
    double* d = nullptr;
    CUDA_TRY(none, cudaMalloc(&d, (size_t)n * 64));
    cudaError_t e = cudaMemcpy(d, io8, (size_t)n * 64, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_kat_urs<<<nblk(n, 128), 128>>>(q, mode, n, d);
        e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess) e = cudaMemcpy(io8, d, (size_t)n * 64, cudaMemcpyDeviceToHost);
    cudaFree(d);
    CUDA_TRY(none, e);
    return 0;
}

int cfd2d_kat_calc_flux(int device, int n, const double* in12, double gam, int flux, double* out4) {
    g_create_error.clear();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); g_create_error = "no usable CUDA device"; return CFD2D_ENODEV; }
    if (n <= 0) return 0;
    cfd2d_fvm* none = nullptr;
    CUDA_TRY(none, cudaSetDevice(device));
    double *din = nullptr, *dout = nullptr;
    CUDA_TRY(none, cudaMalloc(&din, (size_t)n * 96));
    CUDA_TRY(none, cudaMalloc(&dout, (size_t)n * 32));
    CUDA_TRY(none, cudaMemcpy(din, in12, (size_t)n * 96, cudaMemcpyHostToDevice));
    k_kat_flux<<<nblk(n, 128), 128>>>(make_rim(gam), flux, n, din, dout);
    CUDA_TRY(none, cudaDeviceSynchronize());
    CUDA_TRY(none, cudaMemcpy(out4, dout, (size_t)n * 32, cudaMemcpyDeviceToHost));
    cudaFree(din); cudaFree(dout);
    return 0;
}

}  // extern "C"
