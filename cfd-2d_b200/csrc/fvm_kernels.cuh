// fvm_kernels.cuh -- the CUDA kernels of the FVM_TVD path (sm_100a, FP64, HBM-bound gathers).
//
// Data layout in HBM: 32-byte records, one DRAM sector each, so that an
// indirect (neighbour) read costs exactly one sector:
//   U4[c] = {ro, ru, rv, re}                 conservative state (two buffers: Ua = step start /
//                                            result, Ub = after RK stage 1)
//   W4[c] = {r, p, u, v}                     primitive cache, written by the update kernels
//   G8[c] = {Rx,Ry,Px,Py, Ux,Uy,Vx,Vy}       Green-Gauss gradients (two sectors)
//   F4[e] = {fr,fu,fv,fe} * (l/2)            edge flux staging
// plus SoA per-slot / per-edge geometry read fully coalesced.
#pragma once
#include "fvm_device.cuh"
#include "fvm_riemann_fast.cuh"

struct KParams {
    int nc, nc_ex, ne, nmat;
    int order, flux, max_newton, steady;
    double CFL;
    double lim[5];
    RimC rim;
    // tables
    const MatC* mat;        // [nmat]
    const int* bc_kind;     // [nbc]
    const double* bc_par;   // [4*nbc]
    const unsigned char* cell_mat; // [nc_ex]
    // per (cell, slot) gather tables, slot-major [3][nc]
    const int* s_nb;        // neighbour cell, or -1-bc on a boundary edge
    const double* s_nx;     // outward normal = sign * Edge::n
    const double* s_ny;
    const double* s_l;      // Edge::l
    const int* s_es;        // edge*2 + (cell is c2)
    const double* cell_S;   // [nc_ex]
    // per edge
    const int2* e_c;        // {c1, c2}
    const double2* e_n;     // Edge::n
    const double* e_l2;     // Edge::l * 0.5
    const double4* e_d1;    // Gauss points relative to c1's centre {g1x-cx,g1y-cy,g2x-cx,g2y-cy}
    const double4* e_d2;    // same for c2 (unused on boundary edges)
    const int* e_bc;
    // state
    double* cfl;            // cTau/S per owned cell
    double* ctau;           // cTau
    unsigned int* flag;
    int* err;               // [0] Newton-cap hits  [1] flagged-cell count  [2] Newton iterations (KAT)
    int* lim_list;          // compacted flagged cells, CALLER ids (unordered until sorted)
    int lim_cap;                         // >= nc
    unsigned char* rstat;   // [nc_ex] remediation sweep status, all zero between steps
    // cell numbering: the device numbers owned cells along a Hilbert curve (fvm_tiling.h); the
    // caller's numbering only matters for I/O and for the sweep order of remediateLimCells
    const int* c_perm;      // [nc_ex] caller -> device
    const int* c_orig;      // [nc_ex] device -> caller
};

// A record is one 32-byte DRAM sector; sm_100 moves it with ONE 256-bit access (SASS LDG.E.ENL2.256 /
// STG.E.ENL2.256, PTX ld/st.global.v4.f64) instead of two 128-bit ones: half the load/store
// instructions and L1 tag look-ups in every sweep.  Records are 32-byte aligned (cudaMalloc base,
// 32-byte stride).  CFD2D_LD256=0 keeps the two 16-byte accesses.
#ifndef CFD2D_LD256
#define CFD2D_LD256 1
#endif
// L2 fill granularity of a record load: "" = the 32-byte sector asked for; ".L2::128B" / ".L2::256B" ask the
// L2 to fetch the whole line(s) around it (SASS LDG...LTC128B / LTC256B) -- neighbouring records in Hilbert
// order are wanted by neighbouring threads a moment later.  Measured: no effect either way (profiles/r2zm).
#ifndef CFD2D_LD_L2HINT
#define CFD2D_LD_L2HINT ""
#endif
__device__ __forceinline__ double4 ld4(const double4* __restrict__ p, int i) {
#if CFD2D_LD256
    double4 v;
    asm("ld.global.nc" CFD2D_LD_L2HINT ".v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p + i));
    return v;
#else
    const double2* q = reinterpret_cast<const double2*>(p + i);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
#endif
}
__device__ __forceinline__ double4 ld4cg(const double4* p, int i) {
#if CFD2D_LD256
    double4 v;
    asm volatile("ld.global.cg" CFD2D_LD_L2HINT ".v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p + i) : "memory");
    return v;
#else
    const double2* q = reinterpret_cast<const double2*>(p + i);
    double2 a = __ldcg(q), b = __ldcg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
#endif
}
__device__ __forceinline__ void st4(double4* p, int i, double4 v) {
#if CFD2D_LD256
    asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" :: "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w), "l"(p + i) : "memory");
#else
    double2* q = reinterpret_cast<double2*>(p + i);
    q[0] = make_double2(v.x, v.y);
    q[1] = make_double2(v.z, v.w);
#endif
}

// L2 prefetch of one 128-byte line (hint; never faults).  Every block of k_grad asks for the COALESCED
// tables (neighbour ids, slot geometry, areas, own-cell records) of the block CFD2D_GRAD_PF_BLOCKS further
// on, which will run about half a resident wave later: that block's first-level loads are then L2 hits and
// only the gathers they feed still pay a DRAM latency.
// Measured at 4 M cells (profiles/r2zf ... r2zj): k_grad 0.143 -> 0.122 ms with the distance at half a resident wave
// (one wave 0.127, two waves: nothing, the lines are gone again).  The same prefetch was tried and removed for
// k_update (slower: it already runs at 81-92 % of the DRAM peak, the prefetches only add requests), k_flux (no
// gain at 0.25 ... 3 waves: the first-level loads are a small part of its stalls) and k_cell_lf1 (slower).
#ifndef CFD2D_GRAD_PF_BLOCKS
#define CFD2D_GRAD_PF_BLOCKS 888   // 0: off
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ MatC get_mat(const KParams& P, int c) {
    int im = (P.nmat > 1) ? (int)P.cell_mat[c] : 0;
    return P.mat[im];
}

// ---------------------------------------------------------------------------------------------
// K0: U4 -> W4 for cells [c0, c1)  (FVM_TVD::convertConsToPar, fvm_tvd.cpp:803-813)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prim(KParams P, const double4* __restrict__ U, double4* __restrict__ W, int c0, int c1) {
    int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    double4 u = ld4cg(U, c);
    MatC m = get_mat(P, c);
    Prim w = cons_to_prim(u.x, u.y, u.z, u.w, m.gm1);
    st4(W, c, make_double4(w.r, w.p, w.u, w.v));
}

// the same for a list of cells (multi-rank pipe layout: the cells around the send set)
__global__ void __launch_bounds__(256) k_prim_list(KParams P, const double4* __restrict__ U, double4* __restrict__ W,
                                                   const int* __restrict__ list, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = __ldg(list + i);
    double4 u = ld4cg(U, c);
    MatC m = get_mat(P, c);
    Prim w = cons_to_prim(u.x, u.y, u.z, u.w, m.gm1);
    st4(W, c, make_double4(w.r, w.p, w.u, w.v));
}

// ---------------------------------------------------------------------------------------------
// K1: time step (FVM_TVD::calcTimeStep, fvm_tvd.cpp:216-240).
//   steady:   cTau[c] = CFL*S/max(|u|+cz,|v|+cz), cfl[c] = cTau/S
//   unsteady: block/warp-shuffle min of the same quantity -> atomicMin on the (positive) double's
//             bit pattern; min is order independent => bit-safe
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double cell_tau(const KParams& P, const double4& w4, int c) {
    MatC m = get_mat(P, c);
    Prim w = {w4.x, w4.y, w4.z, w4.w};
    double cz = prim_cz(w, m);
    double a = fabs(w.u) + cz, b = fabs(w.v) + cz;
    double mx = (a > b) ? a : b;
    return P.CFL * P.cell_S[c] / mx;
}

__global__ void __launch_bounds__(256) k_tau_steady(KParams P, const double4* __restrict__ W) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.nc) return;
    double t = cell_tau(P, ld4cg(W, c), c);
    P.ctau[c] = t;
    P.cfl[c] = t / P.cell_S[c];
}

__global__ void __launch_bounds__(256) k_tau_min(KParams P, const double4* __restrict__ W, unsigned long long* __restrict__ out_bits) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    double t = __longlong_as_double(0x7ff0000000000000LL);  // +inf
    if (c < P.nc) t = cell_tau(P, ld4cg(W, c), c);
    // the reference's "if (TAU > x) TAU = x" ignores NaN candidates; so does fmin
    for (int o = 16; o > 0; o >>= 1) t = fmin(t, __shfl_xor_sync(0xffffffffu, t, o));
    __shared__ double sm[8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sm[w] = t;
    __syncthreads();
    if (w == 0) {
        t = (l < 8) ? sm[l] : __longlong_as_double(0x7ff0000000000000LL);
        for (int o = 4; o > 0; o >>= 1) t = fmin(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (l == 0 && t > 0.0) atomicMin(out_bits, (unsigned long long)__double_as_longlong(t));
    }
}

__global__ void __launch_bounds__(256) k_tau_fill(KParams P, double tau) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.nc) return;
    P.ctau[c] = tau;
    P.cfl[c] = tau / P.cell_S[c];
}

// ---------------------------------------------------------------------------------------------
// K2: boundary ghost state + Green-Gauss gradient GATHER (FVM_TVD::calcGrad, fvm_tvd.cpp:242-301;
// boundaryCond :694-711).  One thread per owned cell; the three edge terms are added in the cell's
// edgesInd (= ascending edge id) order, each formed exactly as the reference's edge loop forms it
// -- ((pL+pR)/2 * n.x) * l with the edge's own normal, "+=" on the c1 side and "-=" on the c2 side
// (the outward normal s*n reproduces the sign exactly) -- then the true division by S.  No atomics.
// ---------------------------------------------------------------------------------------------
// `list` (optional): only these n cells (multi-GPU: the cells whose gradients are sent to a peer).
#ifndef CFD2D_GRAD_NT
#define CFD2D_GRAD_NT 64       // threads per block of k_grad: 256 -> 0.147, 128 -> 0.144, 64 -> 0.143 ms at 4 M cells
#endif
#ifndef CFD2D_UPDATE_NT
#define CFD2D_UPDATE_NT 256
#endif
#ifndef CFD2D_GRAD_MINB
#define CFD2D_GRAD_MINB (768 / CFD2D_GRAD_NT)     // 85 registers, no spills, 24 warps/SM with every load of a thread in flight at once:
#endif                        // 0.144 ms at 4 M cells; 64 registers (32 warps, spills) 0.151; loads consumed slot by slot 0.162
// `skip_halo_adjacent` (multi-rank interior pass, list == nullptr): leave out the cells with a rank-halo
// neighbour (id >= nc) -- they are the `list` of the boundary pass on the comm stream; the test costs no
// memory traffic, whereas a 4-byte-per-cell interior list made this sweep 9 % slower.
__global__ void __launch_bounds__(CFD2D_GRAD_NT, CFD2D_GRAD_MINB) k_grad(KParams P, const double4* __restrict__ W, double4* __restrict__ G,
                                              const int* __restrict__ list, int n, int skip_halo_adjacent) {
    __shared__ double s_park[CFD2D_GRAD_NT];
    int c = blockIdx.x * blockDim.x + threadIdx.x;
#if CFD2D_GRAD_PF_BLOCKS
    if (!list) {
        constexpr int NT = CFD2D_GRAD_NT, LI = NT * 4 / 128, LD = NT * 8 / 128, LW = NT * 32 / 128;   // lines per table
        const long long cb = ((long long)blockIdx.x + (long long)CFD2D_GRAD_PF_BLOCKS) * NT;
        if (cb + NT <= n) {
            int t = threadIdx.x;
            const char* q = nullptr;
            if (t < 3 * LI) q = (const char*)(P.s_nb + (size_t)(t / LI) * P.nc + cb) + 128 * (t % LI);
            else if ((t -= 3 * LI) < 9 * LD) {
                const int tab = t / LD, k = tab % 3;
                const double* base = tab < 3 ? P.s_nx : (tab < 6 ? P.s_ny : P.s_l);
                q = (const char*)(base + (size_t)k * P.nc + cb) + 128 * (t % LD);
            } else if ((t -= 9 * LD) < LD) q = (const char*)(P.cell_S + cb) + 128 * t;
            else if ((t -= LD) < LW) q = (const char*)(W + cb) + 128 * t;
            if (q) prefetch_l2(q);
        }
    }
#endif
    if (c >= n) return;
    if (list) c = __ldg(list + c);
    // Loads first, arithmetic after: the three neighbour ids, then ALL three neighbour records
    // back-to-back (a boundary slot re-reads the cell's own record: same sector, no branch), so a
    // thread pays one exposed gather latency, not one per slot (ncu: 62 % of the stall samples were
    // long-scoreboard waits spread over four separate points).  The area is needed last; it is loaded
    // with the rest and parked in shared memory so the compiler cannot sink the load next to its use.
    int nb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) nb[k] = __ldg(P.s_nb + (size_t)k * P.nc + c);
    if (skip_halo_adjacent && (nb[0] >= P.nc || nb[1] >= P.nc || nb[2] >= P.nc)) return;
    const double4 ws = ld4(W, c);
    double4 wnb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) wnb[k] = ld4(W, nb[k] >= 0 ? nb[k] : c);
    volatile double* park = s_park + threadIdx.x;
    *park = P.cell_S[c];
    double g[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double gnx[3], gny[3], gl[3];      // the nine slot-geometry loads go out with the gathers, before any arithmetic
#pragma unroll
    for (int k = 0; k < 3; k++) {
        gnx[k] = __ldg(P.s_nx + (size_t)k * P.nc + c);
        gny[k] = __ldg(P.s_ny + (size_t)k * P.nc + c);
        gl[k] = __ldg(P.s_l + (size_t)k * P.nc + c);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double nx = gnx[k], ny = gny[k], l = gl[k];
        double4 wn = wnb[k];
        if (nb[k] < 0) {
            int ib = -1 - nb[k];
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
    double si = *park;
    st4(G, 2 * c, make_double4(g[0] / si, g[1] / si, g[2] / si, g[3] / si));
    st4(G, 2 * c + 1, make_double4(g[4] / si, g[5] / si, g[6] / si, g[7] / si));
}

// ---------------------------------------------------------------------------------------------
// K3: per-edge linear reconstruction at the two Gauss points + numerical flux
// (FVM_TVD::reconstruct fvm_tvd.cpp:646-691, calcFlux :602-643, rim_orig global.cpp:232-405, the
// Gauss-point loop of run() :341-352).  Writes F4[e] = (sum over GPs) * (l*0.5), the quantity the
// reference scatters (:353-363).  FLUX: 0 Godunov with the bit-faithful rim_orig_dev, 1 Lax-Friedrichs,
// 2 Godunov with the reduced-instruction solver (fvm_riemann_fast.cuh; the default for Godunov
// handles).  ORDER: 2 linear, 1 constant.
//
// Work decomposition: ONE THREAD PER (edge, Gauss point); the two Gauss points of an edge sit in
// adjacent lanes.  The exact Riemann solver is a long chain of dependent FP64 divisions/sqrt/exp/log
// (latency bound at the ~12 warps/SM a 2-GP-per-thread version reaches, ncu profiles/r01_*), so
// halving the per-thread state doubles the warps in flight, and lane pairs solve nearly identical
// problems, which halves the number of distinct branch paths per warp.  The pair's fluxes are
// combined with one shuffle: fr = (0.0 + fr1) + fr2 in the reference; IEEE addition commutes, so
// both lanes form the same sum bit for bit; lane 0 stores (fr,fu), lane 1 stores (fv,fe).
// ---------------------------------------------------------------------------------------------
#ifndef CFD2D_FLUX_NT
#define CFD2D_FLUX_NT 64      // threads per block of k_flux: 256 -> 0.383, 128 -> 0.366, 64 -> 0.359, 32 -> 0.357 ms (Godunov, 4 M cells):
                              // small blocks give their warp slots back sooner when Newton trip counts differ
#endif
#ifndef CFD2D_FLUX_MINB
#define CFD2D_FLUX_MINB (1024 / CFD2D_FLUX_NT)     // 64 registers per thread
#endif
#ifndef CFD2D_FLUXLF_MINB
#define CFD2D_FLUXLF_MINB (1024 / CFD2D_FLUX_NT)
#endif

// LANE SPLIT: the pair's lane 0 gathers cell c1, lane 1 gathers cell c2 -- each lane reads ONE
// cell's records (W, both gradient sectors, that cell's Gauss-point offsets `d`) and reconstructs
// that cell's state at BOTH Gauss points; the state at the partner's Gauss point goes to the
// partner by shuffle.  Against "every lane gathers both cells" this issues the gathers of both
// cells at once (one exposed DRAM latency per thread instead of two: ncu put 24 % of the kernel's
// warp-stall samples on the second gather), and halves the gather instructions and L1 look-ups.
// Same expressions on the same operands => same bits.  On a boundary edge both lanes read c1.
//
// flux_gp_tail: everything after the loads.  w/ga/gb/d = this lane's cell records, Eown = its total
// specific energy (LF only), l2slot = where Edge::l*0.5 of this edge is parked in shared memory.
template <int FLUX, int ORDER>
__device__ __forceinline__ void flux_gp_tail(const KParams& P, double4* __restrict__ F, int scale_by_l2, int e, const int gp,
                                             const bool live, const bool inner, const int c1, const double2 n,
                                             const double4 w, const double4 ga, const double4 gb, const double4 d,
                                             const double Eown, const volatile double* l2slot) {
    Prim M = {w.x, w.y, w.z, w.w};      // my cell at my Gauss point
    Prim O = M;                          // my cell at the partner lane's Gauss point
    if (ORDER == 2) {
        const double mx = gp ? d.z : d.x, my = gp ? d.w : d.y;
        const double ox = gp ? d.x : d.z, oy = gp ? d.y : d.w;
        constexpr bool FM = FLUX == 2;
        M.r = recon1<FM>(M.r, ga.x, ga.y, mx, my);
        M.p = recon1<FM>(M.p, ga.z, ga.w, mx, my);
        M.u = recon1<FM>(M.u, gb.x, gb.y, mx, my);
        M.v = recon1<FM>(M.v, gb.z, gb.w, mx, my);
        O.r = recon1<FM>(O.r, ga.x, ga.y, ox, oy);
        O.p = recon1<FM>(O.p, ga.z, ga.w, ox, oy);
        O.u = recon1<FM>(O.u, gb.x, gb.y, ox, oy);
        O.v = recon1<FM>(O.v, gb.z, gb.w, ox, oy);
    }
    Prim X;                              // the partner's cell at my Gauss point
    X.r = __shfl_xor_sync(0xffffffffu, O.r, 1);
    X.p = __shfl_xor_sync(0xffffffffu, O.p, 1);
    X.u = __shfl_xor_sync(0xffffffffu, O.u, 1);
    X.v = __shfl_xor_sync(0xffffffffu, O.v, 1);
    double EX = 0.0;
    if (FLUX == 1) EX = __shfl_xor_sync(0xffffffffu, Eown, 1);
    Prim L, R;
    double EL = 0.0, ER = 0.0;
    if (inner) {
        L.r = gp ? X.r : M.r; L.p = gp ? X.p : M.p; L.u = gp ? X.u : M.u; L.v = gp ? X.v : M.v;
        R.r = gp ? M.r : X.r; R.p = gp ? M.p : X.p; R.u = gp ? M.u : X.u; R.v = gp ? M.v : X.v;
        if (FLUX == 1) { EL = gp ? EX : Eown; ER = gp ? Eown : EX; }
    } else {
        L = M; EL = Eown;
        MatC m = get_mat(P, c1);
        Prim Lc = {w.x, w.y, w.z, w.w};
        const double T1 = prim_T(Lc, m);                       // cell-centre T, before extrapolation
        int ib = __ldg(P.e_bc + e);
        R = ghost_state(L, T1, P.bc_kind[ib], P.bc_par + 4 * ib, n.x, n.y, m, (FLUX == 1) ? &ER : nullptr);
    }
    double f0, f1, f2, f3;
    if (FLUX == 0) {
        int it = flux_godunov_dev(P.rim, P.max_newton, L, R, n.x, n.y, f0, f1, f2, f3);
        if (it < 0 && live) atomicAdd(P.err, 1);
    } else if (FLUX == 2) {
        int it = flux_godunov_fast(P.rim, P.max_newton, L, R, n.x, n.y, f0, f1, f2, f3);
        if (it < 0 && live) atomicAdd(P.err, 1);
    } else {
        flux_lax_dev(P.rim.GAM, L, EL, R, ER, n.x, n.y, f0, f1, f2, f3);
    }
    // this lane keeps two of the four sums: lane 0 (fr,fu), lane 1 (fv,fe)
    double a = gp ? f2 : f0, b = gp ? f3 : f1;       // mine
    double oa = gp ? f0 : f2, ob = gp ? f1 : f3;     // the partner's pair
    double pa = __shfl_xor_sync(0xffffffffu, oa, 1), pb = __shfl_xor_sync(0xffffffffu, ob, 1);
    double sa = gp ? (pa + a) : (a + pa);            // (0.0 + f_gp1) + f_gp2
    double sb = gp ? (pb + b) : (b + pb);
    if (scale_by_l2) {
        double l2 = *l2slot;
        sa = sa * l2; sb = sb * l2;
    }
    if (live) reinterpret_cast<double2*>(F + e)[gp] = make_double2(sa, sb);
}

// K3, direct form: one thread per (edge, Gauss point), all loads issued up front.
template <int FLUX, int ORDER>
__global__ void __launch_bounds__(CFD2D_FLUX_NT, FLUX != 1 ? CFD2D_FLUX_MINB : CFD2D_FLUXLF_MINB)
k_flux(KParams P, const double4* __restrict__ W, const double4* __restrict__ G,
       const double4* __restrict__ Ucur, double4* __restrict__ F, int scale_by_l2, int e0, int e1) {
    // edges [e0, e1) of the device edge order (multi-rank handles: interior edges first, edges that
    // touch a halo cell last, so the halo exchange overlaps the interior sweep)
    __shared__ double s_park[CFD2D_FLUX_NT];
    const int gp = threadIdx.x & 1;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int e = e0 + (t >> 1);
    const bool live = e < e1;
    if (!live) e = e1 - 1;
    const int2 cc = __ldg(P.e_c + e);
    const double2 n = __ldg(P.e_n + e);
    const bool inner = cc.y >= 0;
    const bool second = gp && inner;
    const int cell = second ? cc.y : cc.x;
    double4 d = make_double4(0.0, 0.0, 0.0, 0.0);
    if (ORDER == 2) d = ld4(second ? P.e_d2 : P.e_d1, e);
    // Edge::l * 0.5 is needed only after the solver; loaded NOW and parked in shared memory (a volatile
    // store cannot be sunk), else the compiler moves the load next to its use and the whole DRAM
    // latency is exposed at the end of every thread (11 % of the stall samples).
    volatile double* park = s_park + threadIdx.x;
    if (scale_by_l2) *park = __ldg(P.e_l2 + e);
    const double4 w = ld4(W, cell);
    double4 ga = make_double4(0.0, 0.0, 0.0, 0.0), gb = ga;
    if (ORDER == 2) { ga = ld4(G, 2 * cell); gb = ld4(G, 2 * cell + 1); }
    double Eown = 0.0;
    if (FLUX == 1) { double4 u = ld4(Ucur, cell); Eown = u.w / u.x; }
    flux_gp_tail<FLUX, ORDER>(P, F, scale_by_l2, e, gp, live, inner, cc.x, n, w, ga, gb, d, Eown, park);
}

// ---------------------------------------------------------------------------------------------
// K4 / K5: deterministic residual GATHER + RK stage update, fused.
//   reference: scatter R[c1] -= F*l2, R[c2] += F*l2 over ascending edges (fvm_tvd.cpp:353-363),
//   then U += (cTau/S)*R for unflagged cells (:366-374); stage 2 additionally the half-sum with the
//   old state and the limit tests (:430-447).  Here each cell sums its three staged edge fluxes in
//   edgesInd (ascending id) order starting from 0.0 -- the same additions in the same order, so R is
//   never stored and no float atomics are needed.  The new primitive cache W4 is written as well.
// STAGE 1: Uout(=Ub) = Uin(=Ua) + cfl*R.     STAGE 2: Ua = 0.5*(Ua + (Ub + cfl*R)), flags.
// ---------------------------------------------------------------------------------------------
template <int STAGE>
__global__ void __launch_bounds__(CFD2D_UPDATE_NT) k_update(KParams P, const double4* __restrict__ F, const double4* Uin,
                                                double4* Uout, double4* __restrict__ W) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.nc) return;
    // every load of the thread is issued before the first branch: the flag test used to sit in front of
    // the slot-table loads, a third dependent DRAM latency (flag -> slots -> fluxes) for a path that is
    // almost never taken
    const unsigned int fl = P.flag[c];
    int es[3];
#pragma unroll
    for (int k = 0; k < 3; k++) es[k] = __ldg(P.s_es + (size_t)k * P.nc + c);
    const double cfl = P.cfl[c];
    double4 u = ld4cg(Uin, c);
    double4 uo = make_double4(0.0, 0.0, 0.0, 0.0);
    if (STAGE == 2) uo = ld4cg(Uout, c);     // Ua: the state at step start (ro_old ...)
    double4 f[3];
#pragma unroll
    for (int k = 0; k < 3; k++) f[k] = ld4cg(F, es[k] >> 1);
    if (fl & 2u) {                       // cellIsLim: frozen until remediated (:368, :421, :432)
        if (STAGE == 1) st4(Uout, c, u);
        else {
            int pos = atomicAdd(P.err + 1, 1);
            if (pos < P.lim_cap) P.lim_list[pos] = __ldg(P.c_orig + c);
        }
        return;
    }
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (es[k] & 1) { r0 += f[k].x; r1 += f[k].y; r2 += f[k].z; r3 += f[k].w; }
        else           { r0 -= f[k].x; r1 -= f[k].y; r2 -= f[k].z; r3 -= f[k].w; }
    }
    u.x += cfl * r0; u.y += cfl * r1; u.z += cfl * r2; u.w += cfl * r3;
    MatC m = get_mat(P, c);
    if (STAGE == 2) {
        u.x = 0.5 * (uo.x + u.x); u.y = 0.5 * (uo.y + u.y); u.z = 0.5 * (uo.z + u.z); u.w = 0.5 * (uo.w + u.w);
    }
    Prim w = cons_to_prim(u.x, u.y, u.z, u.w, m.gm1);
    st4(Uout, c, u);
    st4(W, c, make_double4(w.r, w.p, w.u, w.v));
    if (STAGE == 2) {
        bool lim = (w.r < P.lim[0]) | (w.r > P.lim[1]) | (w.p < P.lim[2]) | (w.p > P.lim[3]) |
                   (fabs(w.u) > P.lim[4]) | (fabs(w.v) > P.lim[4]);
        if (lim) {
            P.flag[c] = fl | 2u;         // setCellFlagLim
            int pos = atomicAdd(P.err + 1, 1);
            if (pos < P.lim_cap) P.lim_list[pos] = __ldg(P.c_orig + c);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4c / K5c: the WHOLE RK stage of the first-order Lax-Friedrichs scheme (BASELINE configs[0]) in one
// cell-parallel sweep.  Without reconstruction there is no gradient pass, and the LF flux is cheap
// enough (2 sqrt, a few divisions) to be evaluated from BOTH sides of an edge, so the staged edge
// fluxes (32 B written + 64 B gathered per edge) disappear as well: a cell reads its own and its
// three neighbours' records and writes its new state -- ~220 B/cell/stage instead of ~400.
// Bit-exactness: an edge's flux is formed by flux_lax_dev with the edge's own orientation
// (L = c1, R = c2, Edge::n) whichever cell evaluates it, so both cells get the same bits as k_flux;
// the two Gauss points of an edge see identical states, so the reference's (0.0 + f) + f is f + f;
// the residual is accumulated in Cell::edgesInd order from 0.0 and updated exactly as in k_update.
// W is ping-ponged (neighbours still read the old primitive state).
// ---------------------------------------------------------------------------------------------
#ifndef CFD2D_LF1_MINB
#define CFD2D_LF1_MINB 3     // registers for every load of a thread in flight at once (see k_grad)
#endif
template <int STAGE>
__global__ void __launch_bounds__(256, CFD2D_LF1_MINB) k_cell_lf1(KParams P, const double4* __restrict__ W, const double4* Uin, double4* Uout,
                                                  double4* __restrict__ Wout, const int* __restrict__ list, int n) {
    // `list` (optional): only these n cells (multi-rank: cells without / with a halo neighbour)
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    if (list) c = __ldg(list + c);
    // loads first (neighbour ids, then all neighbour records back-to-back, geometry), arithmetic after
    const unsigned int fl = P.flag[c];
    int nb[3], esb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { nb[k] = __ldg(P.s_nb + (size_t)k * P.nc + c); esb[k] = __ldg(P.s_es + (size_t)k * P.nc + c); }
    const double4 wc = ld4(W, c);
    double4 u = ld4cg(Uin, c);
    double4 uo = make_double4(0.0, 0.0, 0.0, 0.0);
    if (STAGE == 2) uo = ld4cg(Uout, c);     // Ua: the state at step start (ro_old ...)
    const double cfl = P.cfl[c];
    double4 wnb[3];
    double2 unb[3];                          // {rv, re} of the neighbour; its ro is W.r
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int q = nb[k] >= 0 ? nb[k] : c;
        wnb[k] = ld4(W, q);
        unb[k] = __ldcg(reinterpret_cast<const double2*>(Uin + q) + 1);
    }
    double gnx[3], gny[3], gl[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const size_t o = (size_t)k * P.nc + c;
        gnx[k] = __ldg(P.s_nx + o); gny[k] = __ldg(P.s_ny + o); gl[k] = __ldg(P.s_l + o);
    }
    if (fl & 2u) {                       // cellIsLim: frozen until remediated (:368, :421, :432)
        st4(Wout, c, wc);
        if (STAGE == 1) st4(Uout, c, u);
        else {
            int pos = atomicAdd(P.err + 1, 1);
            if (pos < P.lim_cap) P.lim_list[pos] = __ldg(P.c_orig + c);
        }
        return;
    }
    const Prim own = {wc.x, wc.y, wc.z, wc.w};
    const double Eown = u.w / u.x;
    MatC m = get_mat(P, c);
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double onx = gnx[k], ony = gny[k];                         // outward normal = +-Edge::n
        const double l2 = gl[k] * 0.5;                                   // fvm_tvd.cpp:335
        const bool is_c2 = (esb[k] & 1) != 0;
        double f0, f1, f2, f3;
        if (nb[k] >= 0) {
            const double4 wn = wnb[k];
            const Prim oth = {wn.x, wn.y, wn.z, wn.w};
            const double Eoth = unb[k].y / wn.x;                         // re / ro (W.r IS ro)
            if (is_c2) flux_lax_dev(P.rim.GAM, oth, Eoth, own, Eown, -onx, -ony, f0, f1, f2, f3);
            else       flux_lax_dev(P.rim.GAM, own, Eown, oth, Eoth, onx, ony, f0, f1, f2, f3);
        } else {
            const int ib = -1 - nb[k];
            double ER = 0.0;
            Prim R = ghost_state(own, prim_T(own, m), P.bc_kind[ib], P.bc_par + 4 * ib, onx, ony, m, &ER);
            flux_lax_dev(P.rim.GAM, own, Eown, R, ER, onx, ony, f0, f1, f2, f3);
        }
        f0 = (f0 + f0) * l2; f1 = (f1 + f1) * l2; f2 = (f2 + f2) * l2; f3 = (f3 + f3) * l2;   // two Gauss points, then * l/2
        if (is_c2) { r0 += f0; r1 += f1; r2 += f2; r3 += f3; }
        else       { r0 -= f0; r1 -= f1; r2 -= f2; r3 -= f3; }
    }
    u.x += cfl * r0; u.y += cfl * r1; u.z += cfl * r2; u.w += cfl * r3;
    if (STAGE == 2) {
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
            if (pos < P.lim_cap) P.lim_list[pos] = __ldg(P.c_orig + c);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K6: FVM_TVD::remediateLimCells (fvm_tvd.cpp:464-499).  Rare path (returns at once when nothing is
// flagged).  The reference sweeps the cells in ascending id and overwrites in place, so a flagged
// cell c sees, for a flagged neighbour j = edges[].c2,
//     the REMEDIATED value of j when j < c  (already swept),
//     the ORIGINAL   value of j when j > c  (not reached yet), and its own original value when the
// cell itself is the edge's c2 (the reference's self-neighbour quirk, kept).
// That is a dependency DAG, not a serial chain: the flagged cells are snapshotted (Uold), then swept
// as a wavefront by one CTA -- every round computes all pending cells whose lower-numbered flagged
// c2-neighbours are done (reading new values of those, snapshot values of higher-numbered flagged
// ones) and commits them after a barrier.  Meshes from the reference's readers always have c1 < c2
// (an edge is created by its lower-numbered cell, MeshReaderSalomeUnv.cpp:119-217), so there the
// whole sweep is ONE round; chains only arise for caller meshes with other edge orientations.
// Ids are the CALLER's (the reference's sweep order), the sums run in Cell::edgesInd slot order.
// Halo cells (multi-rank) are read as delivered by the state exchange before this kernel: they are
// the owner's pre-remediation values, which is what the serial sweep reads because a halo cell
// that is an edge's c2 has the higher global id (checked at create() when cell_gid is given).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_remediate(KParams P, double4* U, double4* Uold, double4* W, double4* Unew) {
    int n = P.err[1];
    if (n == 0) return;
    if (n > P.lim_cap) n = P.lim_cap;
    const int* a = P.lim_list;                           // caller ids, any order
    unsigned char* rs = P.rstat;                         // 0 not flagged, 1 pending, 3 computed this round, 2 done
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        int c = P.c_perm[a[q]];
        Uold[c] = U[c];
        rs[c] = 1;
    }
    __syncthreads();
    for (;;) {
        int pending = 0;
        for (int q = threadIdx.x; q < n; q += blockDim.x) {
            const int idc = a[q];
            const int c = P.c_perm[idc];
            if (rs[c] != 1) continue;
            int js[3];
            bool ready = true;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                int es = P.s_es[(size_t)k * P.nc + c];
                int j = P.e_c[es >> 1].y;
                js[k] = j;
                if (j >= 0 && j != c && j < P.nc && (rs[j] & 1) && P.c_orig[j] < idc) ready = false;
            }
            if (!ready) { pending = 1; continue; }
            double sRO = 0.0, sRU = 0.0, sRV = 0.0, sRE = 0.0, S = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                int j = js[k];
                if (j >= 0) {
                    double s = P.cell_S[j];
                    // a flagged neighbour the serial sweep has not reached yet: its original value
                    bool later = j != c && j < P.nc && rs[j] != 0 && P.c_orig[j] > idc;
                    double4 u = later ? Uold[j] : U[j];
                    S += s;
                    sRO += u.x * s; sRU += u.y * s; sRV += u.z * s; sRE += u.w * s;
                }
            }
            Unew[q] = make_double4(sRO / S, sRU / S, sRV / S, sRE / S);
            rs[c] = 3;
        }
        __syncthreads();
        for (int q = threadIdx.x; q < n; q += blockDim.x) {
            const int c = P.c_perm[a[q]];
            if (rs[c] != 3) continue;
            double4 u = Unew[q];
            U[c] = u;
            MatC m = get_mat(P, c);
            Prim w = cons_to_prim(u.x, u.y, u.z, u.w, m.gm1);
            W[c] = make_double4(w.r, w.p, w.u, w.v);
            unsigned int fl = P.flag[c];
            fl += 0x010000u;
            if (fl & 0x200000u) fl &= 0x001110u;
            P.flag[c] = fl;
            rs[c] = 2;
        }
        if (!__syncthreads_or(pending)) break;
    }
    for (int q = threadIdx.x; q < n; q += blockDim.x) rs[P.c_perm[a[q]]] = 0;
    if (threadIdx.x == 0) P.err[1] = 0;
}

// ---------------------------------------------------------------------------------------------
// state import/export: the reference's four separate arrays (caller numbering) <-> U4 records
// (device numbering).  One thread per caller cell: the SoA side is coalesced, the record side moves
// whole 32-byte sectors.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_state(int n, const int* __restrict__ perm, const double* __restrict__ ro,
                                                    const double* __restrict__ ru, const double* __restrict__ rv,
                                                    const double* __restrict__ re, double4* __restrict__ U) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st4(U, __ldg(perm + i), make_double4(ro[i], ru[i], rv[i], re[i]));
}

__global__ void __launch_bounds__(256) k_unpack_state(int n, const int* __restrict__ perm, const double4* __restrict__ U,
                                                      double* __restrict__ ro, double* __restrict__ ru,
                                                      double* __restrict__ rv, double* __restrict__ re) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 u = ld4cg(U, __ldg(perm + i));
    ro[i] = u.x; ru[i] = u.y; rv[i] = u.z; re[i] = u.w;
}

// out[i] = in[perm[i]] (device -> caller order) and in[perm[i]] = src[i] (caller -> device order)
template <class T>
__global__ void __launch_bounds__(256) k_gather_perm(int n, const int* __restrict__ perm, const T* __restrict__ in, T* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
template <class T>
__global__ void __launch_bounds__(256) k_scatter_perm(int n, const int* __restrict__ perm, const T* __restrict__ src, T* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[perm[i]] = src[i];
}

// K8: primitive fields for FVM_TVD::save (convertConsToPar per cell, fvm_tvd.cpp:529-572), caller order
__global__ void __launch_bounds__(256) k_primitive_out(KParams P, const double4* __restrict__ U, double* r, double* p, double* T,
                                                       double* u, double* v, double* cz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nc) return;
    int c = __ldg(P.c_perm + i);
    double4 q = ld4cg(U, c);
    MatC m = get_mat(P, c);
    Prim w = cons_to_prim(q.x, q.y, q.z, q.w, m.gm1);
    if (r) r[i] = w.r;
    if (p) p[i] = w.p;
    if (T) T[i] = prim_T(w, m);
    if (u) u[i] = w.u;
    if (v) v[i] = w.v;
    if (cz) cz[i] = prim_cz(w, m);
}

// gradients of caller cell s (two double4 records) -> out8[s][8]
__global__ void __launch_bounds__(256) k_unpack_grad(int n, const int* __restrict__ perm, const double4* __restrict__ G, double* __restrict__ out8) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    int s = i >> 1, k = i & 1;
    double4 g = ld4cg(G, 2 * __ldg(perm + s) + k);
    out8[4 * (size_t)i + 0] = g.x; out8[4 * (size_t)i + 1] = g.y; out8[4 * (size_t)i + 2] = g.z; out8[4 * (size_t)i + 3] = g.w;
}

// ---------------------------------------------------------------------------------------------
// function-level known-answer kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_kat_rim(RimC rc, int max_newton, int fast, int n, const double* __restrict__ in8, double* __restrict__ out5, int* __restrict__ iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* a = in8 + 8 * (size_t)i;
    double RI, EI, PI, UI, VI;
    int it = fast ? rim_orig_fast(rc, max_newton, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], RI, EI, PI, UI, VI)
                  : rim_orig_dev(rc, max_newton, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], RI, EI, PI, UI, VI);
    double* q = out5 + 5 * (size_t)i;
    q[0] = RI; q[1] = EI; q[2] = PI; q[3] = UI; q[4] = VI;
    if (iters) iters[i] = it;
}

// Material::URS (global.cpp:9-30) as the kernels evaluate it: mode 0 is the p / cz part of
// cons_to_prim + prim_cz, mode 1 is prim_T, mode 2 the ghost density of ghost_state.
// io8[n][8] = r,p,e,E,u,v,cz,T in place, like the reference's Param.
__global__ void k_kat_urs(MatC m, int mode, int n, double* __restrict__ io8) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* a = io8 + 8 * (size_t)i;
    if (mode == 0) {
        // cons_to_prim forms e = E - 0.5*(u*u+v*v) itself; feed it conservative variables that give back this e
        Prim w; w.r = a[0]; w.u = 0.0; w.v = 0.0;
        w.p = urs_p(a[0], a[2], m.gm1);
        a[1] = w.p;
        a[6] = prim_cz(w, m);
    } else if (mode == 1) {
        Prim w; w.r = a[0]; w.p = a[1]; w.u = 0.0; w.v = 0.0;
        a[2] = urs_e(a[1], a[0], m.gm1);
        a[7] = prim_T(w, m);
    } else {
        Prim w; w.p = a[1]; w.u = 0.0; w.v = 0.0;
        w.r = urs_r(a[1], a[7], m.M);
        a[0] = w.r;
        a[6] = prim_cz(w, m);
    }
}

__global__ void k_kat_flux(RimC rc, int flux, int n, const double* __restrict__ in12, double* __restrict__ out4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* a = in12 + 12 * (size_t)i;
    Prim L = {a[0], a[1], a[2], a[3]}, R = {a[5], a[6], a[7], a[8]};
    double f0, f1, f2, f3;
    if (flux == 1) flux_lax_dev(rc.GAM, L, a[4], R, a[9], a[10], a[11], f0, f1, f2, f3);
    else if (flux == 2) flux_godunov_fast(rc, 1000, L, R, a[10], a[11], f0, f1, f2, f3);
    else flux_godunov_dev(rc, 1000, L, R, a[10], a[11], f0, f1, f2, f3);
    double* q = out4 + 4 * (size_t)i;
    q[0] = f0; q[1] = f1; q[2] = f2; q[3] = f3;
}
