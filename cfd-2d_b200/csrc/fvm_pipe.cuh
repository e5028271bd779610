// fvm_pipe.cuh -- k_stage_pipe: one RK stage of FVM_TVD::run (fvm_tvd.cpp:323-374 / :376-447: calcGrad
// :242-301, the edge-flux loop :329-365, the cell update :366-374 / :419-447) as ONE persistent,
// software-pipelined sm_100a kernel.
//
// What the three-sweep layout costs (profiles/README.md): gradients (64 B/cell) and edge fluxes
// (32 B/edge) are written to HBM and gathered back, the primitive cache W is written and gathered,
// geometry is stored once per sweep -- 643 B per cell and stage, twice the 320 B inventory.  The
// first tile kernel (fvm_fused.cuh) kept gradients and fluxes on the SM but fed itself with plain
// loads behind four barriers and lost to latency.  This kernel changes how the bytes arrive:
//
//   * persistent CTAs walk the Hilbert-ordered tiles (tile = TC consecutive owned cells);
//   * everything static a tile needs is ONE contiguous blob (fvm_tiling.h, build_pipe_plan): cell
//     centres and areas, per-cell edge slots as 16-bit local ids, the ring-1 gradient tables, and
//     64 B per edge of geometry.  One elected thread moves the blob of tile i+1 and the tile's own
//     contiguous state ranges (U, and for stage 2 the step-start state, plus cTau/S and the flags)
//     with cp.async.bulk into the other half of a two-stage shared-memory ring while all warps
//     compute tile i; completion is tracked by an mbarrier (expect_tx / complete_tx), not by a
//     thread barrier;
//   * the few cell records outside the tile's own range (ring 1, ring 2; with ranks: halo
//     gradients) are fetched by cp.async gathers issued one tile ahead;
//   * the primitive state is recomputed from U in shared memory (convertConsToPar, three divisions
//     per staged cell), so the W cache leaves HBM altogether: a stage reads ~235 B and writes 32 B
//     per cell.
//
// Bit-exactness: every number is produced by the same device function with the same operands in
// the same order as k_grad / k_flux / k_update (fvm_kernels.cuh): the outward normal of a c2-side
// slot is the exact negation of Edge::n, l/2 and the Gauss-point offsets PE - P(cell) are the same
// single IEEE operations the host tables of the sweeps hold, the residual is gathered in
// Cell::edgesInd order from 0.0.  tests/test_gpu_parity.py holds the three layouts to the same bits.
#pragma once
#include "fvm_kernels.cuh"
#include "fvm_tiling.h"

struct PParams {
    const PipeTile* tiles;
    const int* tile_ids;        // optional list of tiles (multi-rank: interior / boundary passes)
    int n_tiles;                // entries of the list (or tiles [0, n_tiles))
    const int* ring;            // gather lists
    const unsigned char* blob;
    const int* c_orig;          // device -> caller cell id (flagged-cell list is kept in caller ids)
    // shared-memory map (bytes), from the plan's maxima
    int stage_bytes;            // one half of the ring
    int so_u, so_uold, so_cfl, so_flag, so_ring, so_gx;   // offsets inside a half (blob at 0)
    int o_stage0;               // first half (mbarriers, tile descriptors and the material table sit in front of it)
    int o_W0, o_W1, o_E, o_G, o_F;                         // computed planes
    int nl2_max, nl_max, ne_max;
};

// ---- sm_90+/sm_100a asynchronous-copy primitives (PTX ISA: mbarrier, cp.async.bulk) -----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One thread: arm the mbarrier with the byte count and launch the bulk copies of one tile
template <int STAGE>
__device__ __forceinline__ void pipe_issue_tile(const KParams& P, const PParams& Q, const PipeTile& ti, const double4* Uin,
                                                const double4* Uold, uint32_t sbase, uint32_t bar) {
    const uint32_t b_blob = (uint32_t)ti.blob_bytes;
    const uint32_t b_u = 32u * (uint32_t)ti.n_own;
    const uint32_t b_cfl = (8u * (uint32_t)ti.n_own + 15u) & ~15u;      // the arrays carry 16 bytes of slack
    const uint32_t b_flag = (4u * (uint32_t)ti.n_own + 15u) & ~15u;
    mbar_arrive_expect_tx(bar, b_blob + b_u + (STAGE == 2 ? b_u : 0u) + b_cfl + b_flag);
    bulk_g2s(sbase, Q.blob + ((size_t)ti.blob_off << 4), b_blob, bar);
    bulk_g2s(sbase + Q.so_u, Uin + ti.cbeg, b_u, bar);
    if (STAGE == 2) bulk_g2s(sbase + Q.so_uold, Uold + ti.cbeg, b_u, bar);
    bulk_g2s(sbase + Q.so_cfl, P.cfl + ti.cbeg, b_cfl, bar);
    bulk_g2s(sbase + Q.so_flag, P.flag + ti.cbeg, b_flag, bar);
}

#define PIPE_HDR_BYTES 640      // [0,16) two mbarriers | [32,272) three PipeTile slots | [288,544) up to 8 materials
#define PIPE_MAT_SMEM 8
#ifndef PIPE_LF_BOTH_GP
#define PIPE_LF_BOTH_GP 0      // 1: LF flux phase = one thread per edge, both Gauss points (measured: see profiles/README.md)
#endif

// reconstruction of one side at one Gauss point (FVM_TVD::reconstruct, fvm_tvd.cpp:646-691)
template <bool FM>
__device__ __forceinline__ Prim pipe_recon(const double2 wa, const double2 wb, const double2 a, const double2 b,
                                           const double2 c, const double2 d, double dx, double dy) {
    Prim q = {wa.x, wa.y, wb.x, wb.y};
    q.r = recon1<FM>(q.r, a.x, a.y, dx, dy);
    q.p = recon1<FM>(q.p, b.x, b.y, dx, dy);
    q.u = recon1<FM>(q.u, c.x, c.y, dx, dy);
    q.v = recon1<FM>(q.v, d.x, d.y, dx, dy);
    return q;
}

// FLUX: 0 Godunov (rim_orig_dev), 1 Lax-Friedrichs, 2 Godunov (reduced-instruction solver).
// Uin: the state this stage starts from; Uout: where the stage result goes (stage 1: Ub; stage 2:
// Ua, which is also the step-start state the half-sum reads).
// Flux phase work split: Godunov -- one thread per (edge, Gauss point), pair combined by a shuffle, as
// k_flux (the Riemann solver is a long dependent chain; halving the per-thread state pays).  Lax-
// Friedrichs -- one thread per edge evaluates both Gauss points: the cell records are read from shared
// memory once and the two independent flux evaluations give the FP64 pipe two chains to interleave.
template <int FLUX, int ORDER, int STAGE, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_stage_pipe(KParams P, PParams Q, const double4* Uin, double4* Uout, const double4* __restrict__ Gx) {
    extern __shared__ __align__(128) unsigned char psmem[];
    unsigned char* const smem = psmem;
    const int tid = threadIdx.x;
    const uint32_t s0 = smem_u32(psmem);
    const uint32_t bar0 = s0, bar1 = s0 + 8;
    int* const tslots = reinterpret_cast<int*>(smem + 32);        // 3 x PipeTile (20 ints each)
    MatC* const smat = reinterpret_cast<MatC*>(smem + 288);
    double2* W0 = reinterpret_cast<double2*>(smem + Q.o_W0);     // {r, p} of all staged cells
    double2* W1 = reinterpret_cast<double2*>(smem + Q.o_W1);     // {u, v}
    double* Es = reinterpret_cast<double*>(smem + Q.o_E);        // LF: total specific energy re/ro
    double2* G0 = reinterpret_cast<double2*>(smem + Q.o_G);      // gradients {Rx,Ry} {Px,Py} {Ux,Uy} {Vx,Vy}
    double2* G1 = G0 + (ORDER == 2 ? Q.nl_max : 0);
    double2* G2 = G1 + (ORDER == 2 ? Q.nl_max : 0);
    double2* G3 = G2 + (ORDER == 2 ? Q.nl_max : 0);
    double2* F0 = reinterpret_cast<double2*>(smem + Q.o_F);      // edge fluxes {fr,fu} {fv,fe} * l/2
    double2* F1 = F0 + Q.ne_max;
    static_assert(sizeof(PipeTile) == 80, "PipeTile is 20 ints");

    const int G = gridDim.x;
    int li = blockIdx.x;                                         // position in the tile list
    if (li >= Q.n_tiles) return;
    auto tile_of = [&](int i) { return Q.tile_ids ? __ldg(Q.tile_ids + i) : i; };
    // tile descriptors travel through a 3-slot ring in shared memory (current, next, the one after)
    auto load_desc = [&](int list_pos, int slot) {
        if (tid < 20 && list_pos < Q.n_tiles)
            tslots[slot * 20 + tid] = __ldg(reinterpret_cast<const int*>(Q.tiles + tile_of(list_pos)) + tid);
    };
    const bool mat_smem = P.nmat <= PIPE_MAT_SMEM;
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (mat_smem && tid >= 32 && tid < 32 + P.nmat) smat[tid - 32] = P.mat[tid - 32];
    load_desc(li, 0);
    load_desc(li + G, 1);
    __syncthreads();

    // prologue: loads of the first tile
    {
        const PipeTile& t0 = *reinterpret_cast<const PipeTile*>(tslots);
        if (tid == 0) pipe_issue_tile<STAGE>(P, Q, t0, Uin, Uout, s0 + Q.o_stage0, bar0);
        const uint32_t sb = s0 + Q.o_stage0;
        const int nring = t0.n_l2 - t0.n_own;
        for (int r = tid; r < nring; r += NT) {
            const int c = __ldg(Q.ring + t0.roff + r);
            const double2* src = reinterpret_cast<const double2*>(Uin + c);
            cp_async16(sb + Q.so_ring + 32u * r, src);
            cp_async16(sb + Q.so_ring + 32u * r + 16u, src + 1);
        }
        if (ORDER == 2)
            for (int r = tid; r < t0.n_l - t0.n_g; r += NT) {
                const int c = __ldg(Q.ring + t0.roff + (t0.n_g - t0.n_own) + r);
                const double2* src = reinterpret_cast<const double2*>(Gx + 2 * (size_t)c);
#pragma unroll
                for (int q = 0; q < 4; q++) cp_async16(sb + Q.so_gx + 64u * r + 16u * q, src + q);
            }
        cp_async_commit();
    }
    uint32_t ph0 = 0, ph1 = 0;                                   // mbarrier phase parities
    int slot_cur = 0;
    for (int iter = 0;; iter++) {
        const int b = iter & 1;
        const uint32_t bar = b ? bar1 : bar0;
        unsigned char* sb = smem + Q.o_stage0 + (size_t)b * Q.stage_bytes;
        const bool has_next = li + G < Q.n_tiles;
        const int slot_next = slot_cur == 2 ? 0 : slot_cur + 1, slot_nn = slot_next == 2 ? 0 : slot_next + 1;
        const PipeTile& ti = *reinterpret_cast<const PipeTile*>(tslots + slot_cur * 20);
        const PipeTile& tn = *reinterpret_cast<const PipeTile*>(tslots + slot_next * 20);
        // ---- wait for this tile's data: my own gathers, then the bulk copies
        cp_async_wait_all();
        {
            const uint32_t par = b ? ph1 : ph0;
            uint32_t spins = 0;
            while (!mbar_try_wait(bar, par)) {
                if (++spins > (1u << 24)) { if (tid == 0) atomicExch(P.err + 3, 0x7ead0000 | (iter & 0xffff)); __trap(); }
            }
            if (b) ph1 ^= 1u; else ph0 ^= 1u;
        }
        __syncthreads();   // S0: everybody's gathers are visible; nobody is still in the previous tile
        // ---- start the next tile's bulk copies into the other half (free since S0)
        const uint32_t sbn = s0 + Q.o_stage0 + (uint32_t)(b ^ 1) * Q.stage_bytes;
        if (has_next && tid == 0) pipe_issue_tile<STAGE>(P, Q, tn, Uin, Uout, sbn, b ? bar0 : bar1);
        load_desc(li + 2 * G, slot_nn);                          // read two iterations from now
        // ids of the records the next tile gathers (consumed after the next barrier: latency hidden)
        int rid0 = -1, rid1 = -1;
        const int nring_n = has_next ? tn.n_l2 - tn.n_own : 0;
        if (tid < nring_n) rid0 = __ldg(Q.ring + tn.roff + tid);
        if (tid + NT < nring_n) rid1 = __ldg(Q.ring + tn.roff + tid + NT);

        // views of this tile's half
        const int n_own = ti.n_own, n_g = ti.n_g, n_l = ti.n_l, n_l2 = ti.n_l2, ne_t = ti.ne_t, cbeg = ti.cbeg;
        const double2* cxy = reinterpret_cast<const double2*>(sb + ti.o_cxy);
        const double* cS = reinterpret_cast<const double*>(sb + ti.o_S);
        const unsigned char* cmat = sb + ti.o_mat;
        const unsigned short* slot = reinterpret_cast<const unsigned short*>(sb + ti.o_slot);
        const int SS = (n_own + 7) & ~7;
        const double2* e_n = reinterpret_cast<const double2*>(sb + ti.o_en);
        const double2* e_l = reinterpret_cast<const double2*>(sb + ti.o_el);
        const double2* e_gp = reinterpret_cast<const double2*>(sb + ti.o_egp);
        const double2* Us = reinterpret_cast<const double2*>(sb + Q.so_u);       // 2 x double2 per cell
        auto mat_of = [&](int l) {
            const int im = (P.nmat > 1) ? (int)cmat[l] : 0;
            return mat_smem ? smat[im] : P.mat[im];
        };

        // ---------------- phase P: primitive state of every staged cell (convertConsToPar) --------
        {
            const double2* Ur = reinterpret_cast<const double2*>(sb + Q.so_ring);
            for (int j = tid; j < n_l2; j += NT) {
                const double2* up = j < n_own ? Us + 2 * j : Ur + 2 * (j - n_own);
                const double2 ua = up[0], ub = up[1];
                const double gm1 = mat_of(j).gm1;
                const Prim w = cons_to_prim(ua.x, ua.y, ub.x, ub.y, gm1);
                W0[j] = make_double2(w.r, w.p);
                W1[j] = make_double2(w.u, w.v);
                if (FLUX == 1) Es[j] = ub.y / ua.x;              // pL.E / pR.E of calcFlux's LF block
            }
        }
        __syncthreads();   // S1

        // ---- gathers of the next tile (ring cells' U; with ranks: halo gradients), one tile ahead
        if (has_next) {
            if (rid0 >= 0) {
                const double2* src = reinterpret_cast<const double2*>(Uin + rid0);
                cp_async16(sbn + Q.so_ring + 32u * tid, src);
                cp_async16(sbn + Q.so_ring + 32u * tid + 16u, src + 1);
            }
            if (rid1 >= 0) {
                const double2* src = reinterpret_cast<const double2*>(Uin + rid1);
                cp_async16(sbn + Q.so_ring + 32u * (tid + NT), src);
                cp_async16(sbn + Q.so_ring + 32u * (tid + NT) + 16u, src + 1);
            }
            for (int r = tid + 2 * NT; r < nring_n; r += NT) {
                const int c = __ldg(Q.ring + tn.roff + r);
                const double2* src = reinterpret_cast<const double2*>(Uin + c);
                cp_async16(sbn + Q.so_ring + 32u * r, src);
                cp_async16(sbn + Q.so_ring + 32u * r + 16u, src + 1);
            }
            if (ORDER == 2)
                for (int r = tid; r < tn.n_l - tn.n_g; r += NT) {
                    const int c = __ldg(Q.ring + tn.roff + (tn.n_g - tn.n_own) + r);
                    const double2* src = reinterpret_cast<const double2*>(Gx + 2 * (size_t)c);
#pragma unroll
                    for (int q = 0; q < 4; q++) cp_async16(sbn + Q.so_gx + 64u * r + 16u * q, src + q);
                }
        }
        cp_async_commit();

        // ---------------- phase G: Green-Gauss gradients (k_grad arithmetic) ----------------------
        if (ORDER == 2) {
            const int R = n_g - n_own;
            const int* gnb = reinterpret_cast<const int*>(sb + ti.o_gnb);
            const double* gn = reinterpret_cast<const double*>(sb + ti.o_gn);
            const double2* Gxs = reinterpret_cast<const double2*>(sb + Q.so_gx);
            for (int j = tid; j < n_l; j += NT) {
                if (j < n_g) {
                    const double2 wa = W0[j], wb = W1[j];
                    const double4 ws = make_double4(wa.x, wa.y, wb.x, wb.y);
                    double g[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        int nb;
                        double nx, ny, l;
                        if (j < n_own) {                 // through the tile's edge tables
                            const int es = slot[k * SS + j];
                            const int q = es >> 1;
                            const double2 n = e_n[q], lc = e_l[q];
                            const unsigned int cl = (unsigned int)__double2loint(lc.y);
                            const int l1 = (int)(cl & 0xffffu), l2 = (int)(cl >> 16);
                            l = lc.x;
                            if (es & 1) { nb = l1; nx = -n.x; ny = -n.y; }        // the cell is the edge's c2
                            else { nb = l2 >= 0xff00 ? -1 - (l2 & 0xff) : l2; nx = n.x; ny = n.y; }
                        } else {                         // ring 1: its own small tables
                            const int r = j - n_own;
                            nb = gnb[k * R + r];
                            nx = gn[(k * 3 + 0) * R + r]; ny = gn[(k * 3 + 1) * R + r]; l = gn[(k * 3 + 2) * R + r];
                        }
                        double4 wn;
                        if (nb >= 0) {
                            const double2 a = W0[nb], bq = W1[nb];
                            wn = make_double4(a.x, a.y, bq.x, bq.y);
                        } else {
                            const int ib = -1 - nb;
                            const MatC m = mat_of(j);
                            const Prim pL = {ws.x, ws.y, ws.z, ws.w};
                            const Prim pR = ghost_state(pL, prim_T(pL, m), P.bc_kind[ib], P.bc_par + 4 * ib, nx, ny, m, nullptr);
                            wn = make_double4(pR.r, pR.p, pR.u, pR.v);
                        }
                        double tr = (ws.x + wn.x) / 2, tp = (ws.y + wn.y) / 2, tu = (ws.z + wn.z) / 2, tv = (ws.w + wn.w) / 2;
                        g[0] += tr * nx * l; g[1] += tr * ny * l;
                        g[2] += tp * nx * l; g[3] += tp * ny * l;
                        g[4] += tu * nx * l; g[5] += tu * ny * l;
                        g[6] += tv * nx * l; g[7] += tv * ny * l;
                    }
                    const double si = cS[j];
                    G0[j] = make_double2(g[0] / si, g[1] / si);
                    G1[j] = make_double2(g[2] / si, g[3] / si);
                    G2[j] = make_double2(g[4] / si, g[5] / si);
                    G3[j] = make_double2(g[6] / si, g[7] / si);
                } else {                                 // rank-halo cell: the owner's gradient
                    const double2* gx = Gxs + 4 * (j - n_g);
                    G0[j] = gx[0]; G1[j] = gx[1]; G2[j] = gx[2]; G3[j] = gx[3];
                }
            }
            __syncthreads();   // S2
        }

        // ---------------- phase F: reconstruction + numerical flux (k_flux arithmetic) ------------
        if (FLUX == 1 && PIPE_LF_BOTH_GP) {
            // one thread per edge, both Gauss points
            for (int q = tid; q < ne_t; q += NT) {
                const double2 n = e_n[q], lc = e_l[q];
                const unsigned int cl = (unsigned int)__double2loint(lc.y);
                const int l1 = (int)(cl & 0xffffu), l2 = (int)(cl >> 16);
                const bool inner = l2 < 0xff00;
                const double2 pa = e_gp[2 * q], pb = e_gp[2 * q + 1];
                const double2 wa = W0[l1], wb = W1[l1];
                const double EL = Es[l1];
                Prim La = {wa.x, wa.y, wb.x, wb.y}, Lb = La, Ra, Rb;
                double ERa = 0.0, ERb = 0.0;
                if (ORDER == 2) {
                    const double2 c1 = cxy[l1];
                    const double2 a = G0[l1], bq = G1[l1], cc = G2[l1], dd = G3[l1];
                    La = pipe_recon<FLUX == 2>(wa, wb, a, bq, cc, dd, pa.x - c1.x, pa.y - c1.y);   // DL = PE - P, fvm_tvd.cpp:661-664
                    Lb = pipe_recon<FLUX == 2>(wa, wb, a, bq, cc, dd, pb.x - c1.x, pb.y - c1.y);
                }
                if (inner) {
                    const double2 va = W0[l2], vb = W1[l2];
                    ERa = ERb = Es[l2];
                    Ra.r = va.x; Ra.p = va.y; Ra.u = vb.x; Ra.v = vb.y; Rb = Ra;
                    if (ORDER == 2) {
                        const double2 c2 = cxy[l2];
                        const double2 a = G0[l2], bq = G1[l2], cc = G2[l2], dd = G3[l2];
                        Ra = pipe_recon<FLUX == 2>(va, vb, a, bq, cc, dd, pa.x - c2.x, pa.y - c2.y);
                        Rb = pipe_recon<FLUX == 2>(va, vb, a, bq, cc, dd, pb.x - c2.x, pb.y - c2.y);
                    }
                } else {
                    const int ib = l2 & 0xff;
                    const MatC m = mat_of(l1);
                    const Prim Lc = {wa.x, wa.y, wb.x, wb.y};
                    const double T1 = prim_T(Lc, m);             // cell-centre T, before extrapolation
                    Ra = ghost_state(La, T1, P.bc_kind[ib], P.bc_par + 4 * ib, n.x, n.y, m, &ERa);
                    Rb = ghost_state(Lb, T1, P.bc_kind[ib], P.bc_par + 4 * ib, n.x, n.y, m, &ERb);
                }
                double a0, a1, a2, a3, b0, b1, b2, b3;
                flux_lax_dev(P.rim.GAM, La, EL, Ra, ERa, n.x, n.y, a0, a1, a2, a3);
                flux_lax_dev(P.rim.GAM, Lb, EL, Rb, ERb, n.x, n.y, b0, b1, b2, b3);
                const double l2h = lc.x * 0.5;                   // Edge::l * 0.5, fvm_tvd.cpp:335
                F0[q] = make_double2((a0 + b0) * l2h, (a1 + b1) * l2h);   // (0.0 + f_gp1) + f_gp2, then * l/2
                F1[q] = make_double2((a2 + b2) * l2h, (a3 + b3) * l2h);
            }
            __syncthreads();   // S3
        } else {
            const int nwork = (2 * ne_t + 31) & ~31;
            for (int w = tid; w < nwork; w += NT) {
                int q = w >> 1;
                const int gp = w & 1;
                const bool live = q < ne_t;
                if (!live) q = ne_t - 1;
                const double2 n = e_n[q], lc = e_l[q];
                const unsigned int cl = (unsigned int)__double2loint(lc.y);
                const int l1 = (int)(cl & 0xffffu), l2 = (int)(cl >> 16);
                const bool inner = l2 < 0xff00;
                const double2 pe = e_gp[2 * q + gp];             // the Gauss point of this lane
                const double2 wa = W0[l1], wb = W1[l1];
                Prim L = {wa.x, wa.y, wb.x, wb.y};
                Prim Rr;
                double EL = 0.0, ER = 0.0;
                if (FLUX == 1) EL = Es[l1];
                double T1 = 0.0;
                MatC m;
                if (!inner) { m = mat_of(l1); T1 = prim_T(L, m); }   // cell-centre T, before extrapolation
                if (ORDER == 2) {
                    const double2 c1 = cxy[l1];
                    L = pipe_recon<FLUX == 2>(wa, wb, G0[l1], G1[l1], G2[l1], G3[l1], pe.x - c1.x, pe.y - c1.y);
                }
                if (inner) {
                    const double2 va = W0[l2], vb = W1[l2];
                    Rr.r = va.x; Rr.p = va.y; Rr.u = vb.x; Rr.v = vb.y;
                    if (FLUX == 1) ER = Es[l2];
                    if (ORDER == 2) {
                        const double2 c2 = cxy[l2];
                        Rr = pipe_recon<FLUX == 2>(va, vb, G0[l2], G1[l2], G2[l2], G3[l2], pe.x - c2.x, pe.y - c2.y);
                    }
                } else {
                    const int ib = l2 & 0xff;
                    Rr = ghost_state(L, T1, P.bc_kind[ib], P.bc_par + 4 * ib, n.x, n.y, m, (FLUX == 1) ? &ER : nullptr);
                }
                double f0, f1, f2, f3;
                int it = 0;
                if (FLUX == 0) it = flux_godunov_dev(P.rim, P.max_newton, L, Rr, n.x, n.y, f0, f1, f2, f3);
                else if (FLUX == 2) it = flux_godunov_fast(P.rim, P.max_newton, L, Rr, n.x, n.y, f0, f1, f2, f3);
                else flux_lax_dev(P.rim.GAM, L, EL, Rr, ER, n.x, n.y, f0, f1, f2, f3);
                // perimeter edges are evaluated by two tiles: a Newton-cap hit counts once per evaluation
                if (it < 0 && live) atomicAdd(P.err, 1);
                double a = gp ? f2 : f0, bq = gp ? f3 : f1;      // mine
                double oa = gp ? f0 : f2, ob = gp ? f1 : f3;     // the partner's pair
                double pa = __shfl_xor_sync(0xffffffffu, oa, 1), pb = __shfl_xor_sync(0xffffffffu, ob, 1);
                double sa = gp ? (pa + a) : (a + pa);            // (0.0 + f_gp1) + f_gp2
                double sb2 = gp ? (pb + bq) : (bq + pb);
                const double l2h = lc.x * 0.5;                   // Edge::l * 0.5, fvm_tvd.cpp:335
                sa = sa * l2h; sb2 = sb2 * l2h;
                if (live) { if (gp) F1[q] = make_double2(sa, sb2); else F0[q] = make_double2(sa, sb2); }
            }
            __syncthreads();   // S3
        }

        // ---------------- phase U: residual gather + RK update (k_update arithmetic) --------------
        {
            const double2* Uo = reinterpret_cast<const double2*>(sb + Q.so_uold);
            const double* cfls = reinterpret_cast<const double*>(sb + Q.so_cfl);
            const unsigned int* flags = reinterpret_cast<const unsigned int*>(sb + Q.so_flag);
            for (int j = tid; j < n_own; j += NT) {
                const int c = cbeg + j;
                const unsigned int fl = flags[j];
                const double2 ua = Us[2 * j], ub = Us[2 * j + 1];
                if (fl & 2u) {               // cellIsLim: frozen until remediated (:368, :421, :432)
                    if (STAGE == 1) st4(Uout, c, make_double4(ua.x, ua.y, ub.x, ub.y));
                    else {
                        int pos = atomicAdd(P.err + 1, 1);
                        if (pos < P.lim_cap) P.lim_list[pos] = __ldg(Q.c_orig + c);
                    }
                    continue;
                }
                double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int es = slot[k * SS + j];
                    const double2 fa = F0[es >> 1], fb = F1[es >> 1];
                    if (es & 1) { r0 += fa.x; r1 += fa.y; r2 += fb.x; r3 += fb.y; }
                    else        { r0 -= fa.x; r1 -= fa.y; r2 -= fb.x; r3 -= fb.y; }
                }
                const double cfl = cfls[j];
                double4 u = make_double4(ua.x, ua.y, ub.x, ub.y);
                u.x += cfl * r0; u.y += cfl * r1; u.z += cfl * r2; u.w += cfl * r3;
                if (STAGE == 2) {
                    const double2 oa = Uo[2 * j], ob = Uo[2 * j + 1];     // the state at step start (ro_old ...)
                    u.x = 0.5 * (oa.x + u.x); u.y = 0.5 * (oa.y + u.y); u.z = 0.5 * (ob.x + u.z); u.w = 0.5 * (ob.y + u.w);
                }
                st4(Uout, c, u);
                if (STAGE == 2) {
                    const double gm1 = mat_of(j).gm1;
                    const Prim w = cons_to_prim(u.x, u.y, u.z, u.w, gm1);
                    bool lim = (w.r < P.lim[0]) | (w.r > P.lim[1]) | (w.p < P.lim[2]) | (w.p > P.lim[3]) |
                               (fabs(w.u) > P.lim[4]) | (fabs(w.v) > P.lim[4]);
                    if (lim) {
                        P.flag[c] = fl | 2u; // setCellFlagLim
                        int pos = atomicAdd(P.err + 1, 1);
                        if (pos < P.lim_cap) P.lim_list[pos] = __ldg(Q.c_orig + c);
                    }
                }
            }
        }
        if (!has_next) break;
        li += G;
        slot_cur = slot_next;
    }
}
