// fvm_fused.cuh -- ONE kernel per RK stage: gradients, edge fluxes, residual gather and stage update
// of a tile of cells by one CTA, with the gradients and the edge fluxes held in shared memory.
//
// Replaces, per stage, the three sweeps of FVM_TVD::run (fvm_tvd.cpp:323-374 / :376-447): calcGrad
// (:242-301), the edge-flux loop (:329-365 with reconstruct :646-691, calcFlux :602-643, rim_orig
// global.cpp:232-405) and the cell update (:366-374, stage 2 :419-447).  The plan (tiles, ring-1
// cells, per-tile tables) is built on the host by fvm_tiling.h.
//
// Why: the unfused path moves ~630 B/cell/stage through HBM (gradients written then gathered, edge
// fluxes written then gathered, two sets of geometry); here gradients and fluxes never leave the SM.
// Every value is produced by the same device function and the same operand order as the unfused
// kernels (fvm_kernels.cuh), so fused and unfused states are bit-identical (tests/test_gpu_parity).
//
// Phases of a CTA (tile t):
//   0. stage the primitive state W of the tile's cells and ring 1 in smem (one independent 32-byte
//      load per thread: the memory-level parallelism of the kernel lives here); LF: also E = re/ro.
//   1. gradients of the tile's cells and of the computable ring-1 cells -> smem (cell-parallel
//      gather in Cell::edgesInd order, exactly k_grad; neighbours inside the tile from smem, ring
//      1/2 from L2); ring-1 cells that are rank-halo cells load the gradient received from their
//      owner.
//   2. edge fluxes: one thread per (edge, Gauss point) as in k_flux; W and gradients from smem by
//      local id, edge tables streamed coalesced; F*(l/2) -> smem.
//   3. per owned cell: gather the three staged fluxes in slot order, RK update, new U and W.
// W is double-buffered (other tiles still read the old W while this one writes the new one).
#pragma once
#include "fvm_kernels.cuh"
#include "fvm_tiling.h"

struct FParams {
    const TileInfo* tiles;
    const int* tile_ids;        // optional list of tiles (multi-GPU: interior / boundary passes)
    const int* ring;
    const int* g_nb; const double* g_nx; const double* g_ny; const double* g_l;
    const int* e_c1; const int* e_c2; const unsigned int* e_cl;
    const double2* e_n; const double* e_l2; const double4* e_d1; const double4* e_d2;
    const int* u_es;
    const int* c_orig;          // device -> caller cell id (flagged-cell list is kept in caller ids)
    int nl_max, ne_max;
};

template <int FLUX, int ORDER, int STAGE, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_stage(KParams P, FParams Q, const double4* __restrict__ W, const double4* Uin, double4* Uout,
        double4* __restrict__ Wout, const double4* __restrict__ Gx) {
    extern __shared__ double2 smem[];
    // all per-cell records as double2 planes (16-byte accesses of consecutive cells hit distinct banks)
    double2* W0 = smem;                                   // {r, p} of the tile's cells + ring 1
    double2* W1 = W0 + Q.nl_max;                          // {u, v}
    double2* G0 = W1 + Q.nl_max;                          // gradients: {Rx,Ry} {Px,Py} {Ux,Uy} {Vx,Vy}
    double2* G1 = G0 + (ORDER == 2 ? Q.nl_max : 0);
    double2* G2 = G1 + (ORDER == 2 ? Q.nl_max : 0);
    double2* G3 = G2 + (ORDER == 2 ? Q.nl_max : 0);
    double2* F0 = G3 + (ORDER == 2 ? Q.nl_max : 0);       // edge fluxes {fr,fu} {fv,fe} * l/2
    double2* F1 = F0 + Q.ne_max;
    double* Es = reinterpret_cast<double*>(F1 + Q.ne_max); // LF only: total specific energy re/ro
    const int t = Q.tile_ids ? Q.tile_ids[blockIdx.x] : (int)blockIdx.x;
    const TileInfo ti = Q.tiles[t];
    const int tid = threadIdx.x;

    // ---------------- phase 0: stage the primitive state of the tile + ring 1 -----------------
    // one independent 32-byte load per thread and pass: this is where the memory parallelism is
    for (int j = tid; j < ti.n_l; j += NT) {
        const int c = j < ti.n_own ? ti.cbeg + j : __ldg(Q.ring + ti.roff + (j - ti.n_own));
        double4 w = ld4(W, c);
        W0[j] = make_double2(w.x, w.y);
        W1[j] = make_double2(w.z, w.w);
        if (FLUX == 1) { double4 u = ld4cg(Uin, c); Es[j] = u.w / u.x; }   // pL.E / pR.E of calcFlux's LF block
    }
    __syncthreads();

    // ---------------- phase 1: Green-Gauss gradients (k_grad arithmetic) ----------------------
    if (ORDER == 2) {
        for (int j = tid; j < ti.n_l; j += NT) {
            if (j < ti.n_g) {
                const int c = j < ti.n_own ? ti.cbeg + j : __ldg(Q.ring + ti.roff + (j - ti.n_own));
                double2 wa = W0[j], wb = W1[j];
                double4 ws = make_double4(wa.x, wa.y, wb.x, wb.y);
                double g[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const size_t o = (size_t)ti.goff + (size_t)k * ti.gstride + j;
                    int nb = __ldg(Q.g_nb + o);
                    double nx = __ldg(Q.g_nx + o), ny = __ldg(Q.g_ny + o), l = __ldg(Q.g_l + o);
                    double4 wn;
                    if (nb >= 0) {
                        const unsigned int ln = (unsigned int)(nb - ti.cbeg);
                        if (ln < (unsigned int)ti.n_own) {           // neighbour inside the tile: staged copy
                            double2 a = W0[ln], b = W1[ln];
                            wn = make_double4(a.x, a.y, b.x, b.y);
                        } else wn = ld4(W, nb);                      // ring 1 / ring 2: L2
                    } else {
                        int ib = -1 - nb;
                        MatC m = get_mat(P, c);
                        Prim pL = {ws.x, ws.y, ws.z, ws.w};
                        Prim pR = ghost_state(pL, prim_T(pL, m), P.bc_kind[ib], P.bc_par + 4 * ib, nx, ny, m, nullptr);
                        wn = make_double4(pR.r, pR.p, pR.u, pR.v);
                    }
                    double tr = (ws.x + wn.x) / 2, tp = (ws.y + wn.y) / 2, tu = (ws.z + wn.z) / 2, tv = (ws.w + wn.w) / 2;
                    g[0] += tr * nx * l; g[1] += tr * ny * l;
                    g[2] += tp * nx * l; g[3] += tp * ny * l;
                    g[4] += tu * nx * l; g[5] += tu * ny * l;
                    g[6] += tv * nx * l; g[7] += tv * ny * l;
                }
                double si = __ldg(P.cell_S + c);
                G0[j] = make_double2(g[0] / si, g[1] / si);
                G1[j] = make_double2(g[2] / si, g[3] / si);
                G2[j] = make_double2(g[4] / si, g[5] / si);
                G3[j] = make_double2(g[6] / si, g[7] / si);
            } else {
                // rank-halo cell: the owner's gradient, received by the halo exchange
                const int c = __ldg(Q.ring + ti.roff + (j - ti.n_own));
                double4 ga = ld4cg(Gx, 2 * c), gb = ld4cg(Gx, 2 * c + 1);
                G0[j] = make_double2(ga.x, ga.y);
                G1[j] = make_double2(ga.z, ga.w);
                G2[j] = make_double2(gb.x, gb.y);
                G3[j] = make_double2(gb.z, gb.w);
            }
        }
        __syncthreads();
    }

    // ---------------- phase 2: reconstruction + numerical flux (k_flux arithmetic) ------------
    // only coalesced streaming loads (the tile's edge tables); cell data come from shared memory
    {
        const int nwork = (2 * ti.ne_t + 31) & ~31;
        for (int w = tid; w < nwork; w += NT) {
            int q = w >> 1;
            const int gp = w & 1;
            const bool live = q < ti.ne_t;
            if (!live) q = ti.ne_t - 1;
            const size_t eo = (size_t)ti.eoff + q;
            const unsigned int cl = __ldg(Q.e_cl + eo);
            const int l1 = (int)(cl & 0xffffu), l2 = (int)(cl >> 16);
            const bool inner = l2 != 0xffff;
            double2 n = __ldg(Q.e_n + eo);
            double2 wa = W0[l1], wb = W1[l1];
            Prim L = {wa.x, wa.y, wb.x, wb.y};
            Prim R;
            double EL = 0.0, ER = 0.0;
            if (FLUX == 1) EL = Es[l1];
            double T1 = 0.0;
            MatC m;
            if (!inner) { m = get_mat(P, ti.cbeg + l1); T1 = prim_T(L, m); }   // a boundary edge's c1 is a tile cell
            if (ORDER == 2) {
                double2 d = __ldg(reinterpret_cast<const double2*>(Q.e_d1 + eo) + gp);
                double2 a = G0[l1], b = G1[l1], cc = G2[l1], dd = G3[l1];
                L.r = recon1<FLUX == 2>(L.r, a.x, a.y, d.x, d.y);
                L.p = recon1<FLUX == 2>(L.p, b.x, b.y, d.x, d.y);
                L.u = recon1<FLUX == 2>(L.u, cc.x, cc.y, d.x, d.y);
                L.v = recon1<FLUX == 2>(L.v, dd.x, dd.y, d.x, d.y);
            }
            if (inner) {
                double2 va = W0[l2], vb = W1[l2];
                R.r = va.x; R.p = va.y; R.u = vb.x; R.v = vb.y;
                if (FLUX == 1) ER = Es[l2];
                if (ORDER == 2) {
                    double2 d = __ldg(reinterpret_cast<const double2*>(Q.e_d2 + eo) + gp);
                    double2 a = G0[l2], b = G1[l2], cc = G2[l2], dd = G3[l2];
                    R.r = recon1<FLUX == 2>(R.r, a.x, a.y, d.x, d.y);
                    R.p = recon1<FLUX == 2>(R.p, b.x, b.y, d.x, d.y);
                    R.u = recon1<FLUX == 2>(R.u, cc.x, cc.y, d.x, d.y);
                    R.v = recon1<FLUX == 2>(R.v, dd.x, dd.y, d.x, d.y);
                }
            } else {
                int ib = -1 - __ldg(Q.e_c2 + eo);
                R = ghost_state(L, T1, P.bc_kind[ib], P.bc_par + 4 * ib, n.x, n.y, m, (FLUX == 1) ? &ER : nullptr);
            }
            double f0, f1, f2, f3;
            if (FLUX == 0) {
                int it = flux_godunov_dev(P.rim, P.max_newton, L, R, n.x, n.y, f0, f1, f2, f3);
                // perimeter edges are evaluated by two tiles: count a Newton-cap hit once per evaluation
                if (it < 0 && live) atomicAdd(P.err, 1);
            } else if (FLUX == 2) {
                int it = flux_godunov_fast(P.rim, P.max_newton, L, R, n.x, n.y, f0, f1, f2, f3);
                if (it < 0 && live) atomicAdd(P.err, 1);
            } else {
                flux_lax_dev(P.rim.GAM, L, EL, R, ER, n.x, n.y, f0, f1, f2, f3);
            }
            double a = gp ? f2 : f0, b = gp ? f3 : f1;       // mine
            double oa = gp ? f0 : f2, ob = gp ? f1 : f3;     // the partner's pair
            double pa = __shfl_xor_sync(0xffffffffu, oa, 1), pb = __shfl_xor_sync(0xffffffffu, ob, 1);
            double sa = gp ? (pa + a) : (a + pa);            // (0.0 + f_gp1) + f_gp2
            double sb = gp ? (pb + b) : (b + pb);
            double l2h = __ldg(Q.e_l2 + eo);
            sa = sa * l2h; sb = sb * l2h;
            if (live) { if (gp) F1[q] = make_double2(sa, sb); else F0[q] = make_double2(sa, sb); }
        }
        __syncthreads();
    }

    // ---------------- phase 3: residual gather + RK update (k_update arithmetic) --------------
    for (int j = tid; j < ti.n_own; j += NT) {
        const int c = ti.cbeg + j;
        unsigned int fl = P.flag[c];
        if (fl & 2u) {                       // cellIsLim: frozen until remediated (:368, :421, :432)
            if (STAGE == 1) {
                double4 u = ld4cg(Uin, c);
                st4(Uout, c, u);
                st4(Wout, c, ld4(W, c));
            } else {
                st4(Wout, c, ld4(W, c));     // Wout must describe Uout (= Ua, untouched) -- see below
                int pos = atomicAdd(P.err + 1, 1);
                if (pos < P.lim_cap) P.lim_list[pos] = __ldg(Q.c_orig + c);
            }
            continue;
        }
        double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int es = __ldg(Q.u_es + (size_t)k * P.nc + c);
            double2 fa = F0[es >> 1], fb = F1[es >> 1];
            if (es & 1) { r0 += fa.x; r1 += fa.y; r2 += fb.x; r3 += fb.y; }
            else        { r0 -= fa.x; r1 -= fa.y; r2 -= fb.x; r3 -= fb.y; }
        }
        double cfl = P.cfl[c];
        double4 u = ld4cg(Uin, c);
        u.x += cfl * r0; u.y += cfl * r1; u.z += cfl * r2; u.w += cfl * r3;
        MatC m = get_mat(P, c);
        if (STAGE == 2) {
            double4 uo = ld4cg(Uout, c);     // Ua: the state at step start (ro_old ...)
            u.x = 0.5 * (uo.x + u.x); u.y = 0.5 * (uo.y + u.y); u.z = 0.5 * (uo.z + u.z); u.w = 0.5 * (uo.w + u.w);
        }
        Prim w = cons_to_prim(u.x, u.y, u.z, u.w, m.gm1);
        st4(Uout, c, u);
        st4(Wout, c, make_double4(w.r, w.p, w.u, w.v));
        if (STAGE == 2) {
            bool lim = (w.r < P.lim[0]) | (w.r > P.lim[1]) | (w.p < P.lim[2]) | (w.p > P.lim[3]) |
                       (fabs(w.u) > P.lim[4]) | (fabs(w.v) > P.lim[4]);
            if (lim) {
                P.flag[c] = fl | 2u;         // setCellFlagLim
                int pos = atomicAdd(P.err + 1, 1);
                if (pos < P.lim_cap) P.lim_list[pos] = __ldg(Q.c_orig + c);
            }
        }
    }
}
