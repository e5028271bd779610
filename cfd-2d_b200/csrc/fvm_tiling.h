// fvm_tiling.h -- host-side plan for the tile-fused RK-stage kernel (fvm_fused.cuh).  Pure C++, no
// CUDA: built once per handle in cfd2d_fvm_create(), checked on CPU by tests/test_tiling.py through
// cfd2d_tiling_plan().
//
// The reference sweeps the whole mesh three times per RK stage (calcGrad, the edge-flux loop, the
// cell update: fvm_tvd.cpp:242-301, :329-365, :366-374), and every sweep streams its inputs and
// outputs through memory.  On the GPU the gradients (64 B/cell) and the edge fluxes (32 B/edge)
// never need to leave the SM if one CTA owns a compact TILE of cells:
//   * owned cells of the device are renumbered along a Hilbert curve through the cell centres, a
//     tile is TC consecutive cells of that order (a compact blob with a short perimeter);
//   * ring 1 of a tile = the cells outside it that share an edge with it.  Their gradients are
//     needed by the tile's perimeter edges and are recomputed inside the tile (they in turn read
//     the primitive state of ring 2 straight from HBM/L2) -- except ring-1 cells that are HALO
//     cells of the rank, whose gradients arrive through the NCCL exchange as in the reference
//     (fvm_tvd.cpp: exchange of gradR.. is implied by Method::exchange, method.h:13-127);
//   * the tile's edge list = every edge touching one of its cells; perimeter edges are therefore
//     evaluated by both neighbouring tiles with bit-identical inputs and expression => both get the
//     same bits, no flux exchange, no atomics.
// Renumbering cells does not touch any summation order: per cell, the three edge terms are still
// added in the caller's Cell::edgesInd order (slot order), see fvm_kernels.cuh.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct TileInfo {           // 48 bytes, read uniformly by the CTA
    int cbeg;               // first owned cell (device numbering) of the tile
    int n_own;              // owned cells
    int n_g;                // cells whose gradient is computed in the tile: owned + computable ring 1
    int n_l;                // all local cells: n_g + ring-1 cells that are rank-halo cells
    int roff;               // offset of the ring-1 id list (n_l - n_own entries)
    int goff;               // offset of the gradient gather tables
    int gstride;            // slot stride inside those tables (>= n_g)
    int eoff;               // offset of the edge tables
    int ne_t;               // edges of the tile
    int halo_dep;           // 1 if the tile reads any rank-halo data (ring 1 or ring 2)
    int pad0, pad1;
};

struct HostMesh {           // the caller's cfd2d_mesh after cell renumbering (device numbering)
    int nc = 0, nc_ex = 0, ne = 0;
    std::vector<double> cell_S, cell_cx, cell_cy;
    std::vector<int> cell_mat, cell_edges, edge_c1, edge_c2, edge_bc;
    std::vector<double> edge_nx, edge_ny, edge_l, edge_gp;
};

struct TilePlan {
    int TC = 0, ntiles = 0, nl_max = 0, ne_max = 0;
    std::vector<TileInfo> tiles;
    std::vector<int> ring;                       // ring-1 cell ids per tile
    std::vector<int> g_nb;                       // [goff + k*gstride + j] neighbour id or -1-bc
    std::vector<double> g_nx, g_ny, g_l;         // outward normal (sign*Edge::n), Edge::l
    std::vector<int> e_c1, e_c2;                 // global (device) cell ids; c2 = -1-bc on a boundary edge
    std::vector<uint32_t> e_cl;                  // local ids: l1 | l2 << 16 (l2 = 0xffff on a boundary edge)
    std::vector<int> e_id;                       // caller's edge id (tests / debugging)
    std::vector<int> u_es;                       // [k*nc + c] local edge position*2 + (cell is c2)
    std::vector<int> interior, boundary;         // tile ids without / with rank-halo dependence
    long long sum_ng = 0, sum_ne = 0, sum_ring = 0;
};

// ---- Hilbert curve index of a point on a 2^16 x 2^16 grid
static inline uint32_t hilbert_d(uint32_t x, uint32_t y) {
    uint32_t d = 0;
    for (uint32_t s = 1u << 15; s > 0; s >>= 1) {
        uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = 65535u - x; y = 65535u - y; }
            uint32_t t = x; x = y; y = t;
        }
    }
    return d;
}

// caller -> device cell numbering: owned cells along the Hilbert curve (ties broken by caller id),
// halo cells keep their place (the halo slice layout is the exchange's contract).
static inline void hilbert_cell_order(int nc, int nc_ex, const double* cx, const double* cy, bool enable,
                                      std::vector<int>& perm, std::vector<int>& orig) {
    perm.resize(nc_ex); orig.resize(nc_ex);
    for (int i = 0; i < nc_ex; i++) { perm[i] = i; orig[i] = i; }
    if (!enable || nc < 2) return;
    double x0 = cx[0], x1 = cx[0], y0 = cy[0], y1 = cy[0];
    for (int i = 1; i < nc; i++) {
        x0 = std::min(x0, cx[i]); x1 = std::max(x1, cx[i]);
        y0 = std::min(y0, cy[i]); y1 = std::max(y1, cy[i]);
    }
    double ext = std::max(x1 - x0, y1 - y0);
    if (!(ext > 0.0) || !std::isfinite(ext)) return;
    const double sc = 65535.0 / ext;               // one scale for both axes: tiles stay isotropic
    std::vector<uint64_t> key(nc);
    for (int i = 0; i < nc; i++) {
        double fx = (cx[i] - x0) * sc, fy = (cy[i] - y0) * sc;
        uint32_t ix = fx > 0 ? (fx < 65535.0 ? (uint32_t)fx : 65535u) : 0u;
        uint32_t iy = fy > 0 ? (fy < 65535.0 ? (uint32_t)fy : 65535u) : 0u;
        key[i] = ((uint64_t)hilbert_d(ix, iy) << 32) | (uint32_t)i;
    }
    std::sort(key.begin(), key.end());
    for (int d = 0; d < nc; d++) { int c = (int)(key[d] & 0xffffffffu); orig[d] = c; perm[c] = d; }
}

// the caller's mesh with cells renumbered by perm (edge ids and per-cell slot order unchanged)
template <class M>
static inline void permute_mesh(const M* m, const std::vector<int>& perm, const std::vector<int>& orig, HostMesh& o) {
    const int nc = m->nc, nc_ex = m->nc_ex, ne = m->ne;
    o.nc = nc; o.nc_ex = nc_ex; o.ne = ne;
    o.cell_S.resize(nc_ex); o.cell_cx.resize(nc_ex); o.cell_cy.resize(nc_ex); o.cell_mat.resize(nc_ex);
    for (int d = 0; d < nc_ex; d++) {
        int c = orig[d];
        o.cell_S[d] = m->cell_S[c]; o.cell_cx[d] = m->cell_cx[c]; o.cell_cy[d] = m->cell_cy[c]; o.cell_mat[d] = m->cell_mat[c];
    }
    o.cell_edges.resize(3 * (size_t)nc);
    for (int d = 0; d < nc; d++)
        for (int k = 0; k < 3; k++) o.cell_edges[3 * (size_t)d + k] = m->cell_edges[3 * (size_t)orig[d] + k];
    o.edge_c1.resize(ne); o.edge_c2.resize(ne); o.edge_bc.assign(m->edge_bc, m->edge_bc + ne);
    for (int e = 0; e < ne; e++) {
        o.edge_c1[e] = perm[m->edge_c1[e]];
        o.edge_c2[e] = m->edge_c2[e] >= 0 ? perm[m->edge_c2[e]] : -1;
    }
    o.edge_nx.assign(m->edge_nx, m->edge_nx + ne); o.edge_ny.assign(m->edge_ny, m->edge_ny + ne);
    o.edge_l.assign(m->edge_l, m->edge_l + ne); o.edge_gp.assign(m->edge_gp, m->edge_gp + 4 * (size_t)ne);
}

// direction bin of an edge normal (n and -n folded together), NBIN = boundary edges
static inline int edge_dir_bin(double nx, double ny, int NBIN) {
    const double PI_ = 3.14159265358979323846;
    double a = atan2(ny, nx);
    if (a < 0) a += PI_;
    int b = (int)floor((a + PI_ / (2 * NBIN)) / (PI_ / NBIN));
    if (b >= NBIN || b < 0) b = 0;
    return b;
}

// Build the tile plan on the device-numbered mesh.  Returns "" or an error text.
static inline std::string build_tile_plan(const HostMesh& m, int TC, TilePlan& p) {
    const int nc = m.nc, nc_ex = m.nc_ex, ne = m.ne;
    if (TC < 32) TC = 32;
    if (TC > 16384) TC = 16384;
    p = TilePlan();
    p.TC = TC;
    p.ntiles = (nc + TC - 1) / TC;
    p.tiles.resize(p.ntiles);
    p.u_es.assign(3 * (size_t)nc, 0);
    const int NBIN = 16;
    std::vector<int> cstamp(nc_ex, 0), lidx(nc_ex, 0), estamp(ne, 0), epos(ne, 0);
    std::vector<unsigned char> ebin(ne);
    for (int e = 0; e < ne; e++) ebin[e] = (unsigned char)(m.edge_c2[e] < 0 ? NBIN : edge_dir_bin(m.edge_nx[e], m.edge_ny[e], NBIN));
    std::vector<int> ring_c, ring_h, elist;
    std::vector<std::pair<int, int>> ekey;
    auto other = [&](int e, int c) { return m.edge_c1[e] == c ? m.edge_c2[e] : m.edge_c1[e]; };
    for (int t = 0; t < p.ntiles; t++) {
        TileInfo& ti = p.tiles[t];
        const int cbeg = t * TC, n_own = std::min(TC, nc - cbeg), stamp = t + 1;
        ti.cbeg = cbeg; ti.n_own = n_own; ti.halo_dep = 0; ti.pad0 = ti.pad1 = 0;
        // ---- ring 1 (computable cells first, rank-halo cells last) and the tile's edges
        ring_c.clear(); ring_h.clear(); elist.clear();
        for (int c = cbeg; c < cbeg + n_own; c++) {
            for (int k = 0; k < 3; k++) {
                int e = m.cell_edges[3 * (size_t)c + k];
                if (e < 0 || e >= ne) return "cell_edges entry out of range";
                if (m.edge_c1[e] != c && m.edge_c2[e] != c) return "cell_edges names an edge that does not touch the cell";
                if (estamp[e] != stamp) { estamp[e] = stamp; elist.push_back(e); }
                int nb = other(e, c);
                if (nb >= 0 && (nb < cbeg || nb >= cbeg + n_own) && cstamp[nb] != stamp) {
                    cstamp[nb] = stamp;
                    (nb < nc ? ring_c : ring_h).push_back(nb);
                }
            }
        }
        ti.n_g = n_own + (int)ring_c.size();
        ti.n_l = ti.n_g + (int)ring_h.size();
        if (ti.n_l > 65534) return "tile too large for 16-bit local cell ids";
        if (!ring_h.empty()) ti.halo_dep = 1;
        ti.roff = (int)p.ring.size();
        for (size_t i = 0; i < ring_c.size(); i++) { lidx[ring_c[i]] = n_own + (int)i; p.ring.push_back(ring_c[i]); }
        for (size_t i = 0; i < ring_h.size(); i++) { lidx[ring_h[i]] = ti.n_g + (int)i; p.ring.push_back(ring_h[i]); }
        // ---- gradient gather tables for the n_g cells
        ti.gstride = (ti.n_g + 3) & ~3;
        ti.goff = (int)p.g_nb.size();
        if ((long long)p.g_nb.size() + 3LL * ti.gstride > 0x7fffffffLL) return "mesh too large for 32-bit table offsets";
        p.g_nb.resize(p.g_nb.size() + 3 * (size_t)ti.gstride, -1);
        p.g_nx.resize(p.g_nb.size(), 0.0); p.g_ny.resize(p.g_nb.size(), 0.0); p.g_l.resize(p.g_nb.size(), 0.0);
        for (int j = 0; j < ti.n_g; j++) {
            const int c = j < n_own ? cbeg + j : ring_c[j - n_own];
            for (int k = 0; k < 3; k++) {
                const int e = m.cell_edges[3 * (size_t)c + k];
                if (e < 0 || e >= ne) return "cell_edges entry out of range";
                const size_t o = (size_t)ti.goff + (size_t)k * ti.gstride + j;
                const bool is1 = m.edge_c1[e] == c;
                if (!is1 && m.edge_c2[e] != c) return "cell_edges names an edge that does not touch the cell";
                const int nb = other(e, c);
                p.g_nb[o] = nb >= 0 ? nb : -1 - m.edge_bc[e];
                p.g_nx[o] = is1 ? m.edge_nx[e] : -m.edge_nx[e];
                p.g_ny[o] = is1 ? m.edge_ny[e] : -m.edge_ny[e];
                p.g_l[o] = m.edge_l[e];
                if (nb >= nc) ti.halo_dep = 1;           // ring 2 (or ring 1) in the rank halo
            }
        }
        // ---- edge tables: grouped by normal direction (branch coherence of rim_orig), boundary last
        ekey.resize(elist.size());
        for (size_t i = 0; i < elist.size(); i++) ekey[i] = std::make_pair((int)ebin[elist[i]], elist[i]);
        std::sort(ekey.begin(), ekey.end());
        ti.eoff = (int)p.e_c1.size();
        ti.ne_t = (int)ekey.size();
        for (int q = 0; q < ti.ne_t; q++) {
            const int e = ekey[q].second;
            epos[e] = q;
            const int c1 = m.edge_c1[e], c2 = m.edge_c2[e];
            const bool own1 = c1 >= cbeg && c1 < cbeg + n_own;
            const int l1 = own1 ? c1 - cbeg : lidx[c1];
            int l2 = 0xffff;                                  // boundary edge marker
            if (c2 >= 0) l2 = (c2 >= cbeg && c2 < cbeg + n_own) ? c2 - cbeg : lidx[c2];
            if (c2 < 0 && !own1) return "internal: boundary edge whose cell is outside the tile";
            if (!own1 && cstamp[c1] != stamp) return "internal: edge cell outside tile + ring 1";
            p.e_c1.push_back(c1);
            p.e_c2.push_back(c2 >= 0 ? c2 : -1 - m.edge_bc[e]);
            p.e_cl.push_back((uint32_t)l1 | ((uint32_t)l2 << 16));
            p.e_id.push_back(e);
        }
        for (int c = cbeg; c < cbeg + n_own; c++)
            for (int k = 0; k < 3; k++) {
                const int e = m.cell_edges[3 * (size_t)c + k];
                p.u_es[(size_t)k * nc + c] = epos[e] * 2 + (m.edge_c2[e] == c ? 1 : 0);
            }
        p.nl_max = std::max(p.nl_max, ti.n_l);
        p.ne_max = std::max(p.ne_max, ti.ne_t);
        p.sum_ng += ti.n_g; p.sum_ne += ti.ne_t; p.sum_ring += ti.n_l - n_own;
        (ti.halo_dep ? p.boundary : p.interior).push_back(t);
    }
    return "";
}

// ================================================================================================
// Plan of the PIPELINED tile kernel (fvm_pipe.cuh): the same tiles, but every static table a tile
// needs is packed into ONE contiguous, 16-byte aligned blob so that a single cp.async.bulk moves it
// into shared memory, and the cell records a tile reads from outside its own (contiguous) cell range
// are listed once: ring 1 (computable), ring 1 in the rank halo, ring 2 (state only, feeds the
// gradients of ring 1).
// ================================================================================================
struct PipeTile {           // 80 bytes
    int cbeg, n_own;        // owned cells [cbeg, cbeg + n_own) of the device numbering
    int n_g;                // own + ring-1 cells whose gradient is computed in the tile
    int n_l;                // + ring-1 cells in the rank halo (gradient received from the owner)
    int n_l2;               // + ring-2 cells (state only)
    int ne_t;               // edges touching an owned cell
    int roff;               // offset of the gather list (n_l2 - n_own cell ids) in PipePlan::ring
    int blob_off;           // offset of the blob in 16-byte units
    int blob_bytes;         // multiple of 16
    // byte offsets inside the blob
    int o_cxy, o_S, o_mat, o_slot, o_gnb, o_gn, o_en, o_el, o_egp;
    int halo_dep;
    int pad[1];
};

struct PipePlan {
    int TC = 0, ntiles = 0;
    int nl2_max = 0, nl_max = 0, ne_max = 0, nring_max = 0, nhalo_max = 0, blob_max = 0, nrc_max = 0;
    std::vector<PipeTile> tiles;
    std::vector<int> ring;                    // gather lists
    std::vector<unsigned char> blob;          // all tiles' tables
    std::vector<int> interior, boundary;
    long long sum_ring1 = 0, sum_ring2 = 0, sum_ne = 0;
};

static inline int pad16(long long x) { return (int)((x + 15) & ~15LL); }

// Blob layout of one tile (S = slot stride = n_own rounded up to 8, R = n_g - n_own):
//   o_cxy : double2 {cx, cy}        [n_l]          cell centres (reconstruction: DL = PE - P, fvm_tvd.cpp:661-664)
//   o_S   : double  S               [n_g]          cell areas (gradient division, fvm_tvd.cpp:297-300)
//   o_mat : uint8   material        [n_l2]
//   o_slot: uint16  edge*2+side     [3][S]         Cell::edgesInd order of the owned cells (local edge ids)
//   o_gnb : int32   neighbour       [3][R]         ring-1 gradient gather: local cell id, or -1-bc
//   o_gn  : double  nx, ny, l       [3][3][R]      outward normal and length of those edges
//   o_en  : double2 {nx, ny}        [ne_t]         Edge::n
//   o_el  : {double l; uint32 l1 | l2 << 16; uint32 0} [ne_t]   Edge::l and the local cell ids of the edge
//                                                  (l2 = 0xff00 | bc on a boundary edge)
//   o_egp : double4 Edge::c[1..2]   [ne_t]         the two Gauss points          (64 B per edge in all)
// Planes, not records: lane pairs (one per Gauss point) of consecutive edges then read consecutive
// 16-byte chunks of shared memory.
static inline std::string build_pipe_plan(const HostMesh& m, int TC, bool dir_bins, PipePlan& p) {
    const int nc = m.nc, nc_ex = m.nc_ex, ne = m.ne;
    if (TC < 32) TC = 32;
    if (TC > 8192) TC = 8192;
    TC = (TC + 7) & ~7;                       // bulk copies of the cfl / flag ranges start 16-byte aligned
    p = PipePlan();
    p.TC = TC;
    p.ntiles = (nc + TC - 1) / TC;
    p.tiles.resize(p.ntiles);
    const int NBIN = 16;
    std::vector<int> cstamp(nc_ex, 0), lidx(nc_ex, 0), estamp(ne, 0), epos(ne, 0);
    std::vector<int> r1c, r1h, r2, elist;
    std::vector<std::pair<long long, int>> ekey;
    auto other = [&](int e, int c) { return m.edge_c1[e] == c ? m.edge_c2[e] : m.edge_c1[e]; };
    for (int t = 0; t < p.ntiles; t++) {
        PipeTile& ti = p.tiles[t];
        const int cbeg = t * TC, n_own = std::min(TC, nc - cbeg), stamp = t + 1;
        ti = PipeTile();
        ti.cbeg = cbeg; ti.n_own = n_own;
        r1c.clear(); r1h.clear(); r2.clear(); elist.clear();
        auto in_tile = [&](int c) { return c >= cbeg && c < cbeg + n_own; };
        for (int c = cbeg; c < cbeg + n_own; c++)
            for (int k = 0; k < 3; k++) {
                const int e = m.cell_edges[3 * (size_t)c + k];
                if (e < 0 || e >= ne) return "cell_edges entry out of range";
                if (m.edge_c1[e] != c && m.edge_c2[e] != c) return "cell_edges names an edge that does not touch the cell";
                if (estamp[e] != stamp) { estamp[e] = stamp; elist.push_back(e); }
                const int nb = other(e, c);
                if (nb >= 0 && !in_tile(nb) && cstamp[nb] != stamp) { cstamp[nb] = stamp; (nb < nc ? r1c : r1h).push_back(nb); }
            }
        for (size_t i = 0; i < r1c.size(); i++) {
            const int c = r1c[i];
            for (int k = 0; k < 3; k++) {
                const int e = m.cell_edges[3 * (size_t)c + k];
                if (e < 0 || e >= ne) return "cell_edges entry out of range";
                const int nb = other(e, c);
                if (nb >= 0 && !in_tile(nb) && cstamp[nb] != stamp) { cstamp[nb] = stamp; r2.push_back(nb); }
                if (nb >= nc) ti.halo_dep = 1;
            }
        }
        if (!r1h.empty()) ti.halo_dep = 1;
        ti.n_g = n_own + (int)r1c.size();
        ti.n_l = ti.n_g + (int)r1h.size();
        ti.n_l2 = ti.n_l + (int)r2.size();
        if (ti.n_l2 >= 0xff00) return "tile too large for 16-bit local cell ids";
        ti.roff = (int)p.ring.size();
        for (size_t i = 0; i < r1c.size(); i++) { lidx[r1c[i]] = n_own + (int)i; p.ring.push_back(r1c[i]); }
        for (size_t i = 0; i < r1h.size(); i++) { lidx[r1h[i]] = ti.n_g + (int)i; p.ring.push_back(r1h[i]); }
        for (size_t i = 0; i < r2.size(); i++) { lidx[r2[i]] = ti.n_l + (int)i; p.ring.push_back(r2[i]); }
        auto lid = [&](int c) { return in_tile(c) ? c - cbeg : lidx[c]; };
        auto gid = [&](int l) { return l < n_own ? cbeg + l : p.ring[ti.roff + l - n_own]; };
        // ---- edge order: by normal direction (branch coherence of rim_orig) or by the lower local cell id
        // (shared-memory locality) -- free, F is addressed through the slot table
        ekey.resize(elist.size());
        for (size_t i = 0; i < elist.size(); i++) {
            const int e = elist[i];
            const int c1 = m.edge_c1[e], c2 = m.edge_c2[e];
            const unsigned lo = (unsigned)std::min(lid(c1), c2 >= 0 ? lid(c2) : 0x7fffffff);
            long long key;
            if (dir_bins) key = ((long long)(c2 < 0 ? NBIN : edge_dir_bin(m.edge_nx[e], m.edge_ny[e], NBIN)) << 32) | lo;
            else key = ((long long)(c2 < 0 ? 1 : 0) << 40) | lo;
            ekey[i] = std::make_pair(key, e);
        }
        std::sort(ekey.begin(), ekey.end());
        ti.ne_t = (int)ekey.size();
        if (ti.ne_t >= 32768) return "tile too large for 16-bit local edge ids";
        for (int q = 0; q < ti.ne_t; q++) epos[ekey[q].second] = q;
        // ---- blob
        const int S = (n_own + 7) & ~7, R = ti.n_g - n_own;
        long long o = 0;
        ti.o_cxy = (int)o; o += pad16(16LL * ti.n_l);
        ti.o_S = (int)o; o += pad16(8LL * ti.n_g);
        ti.o_mat = (int)o; o += pad16(ti.n_l2);
        ti.o_slot = (int)o; o += pad16(2LL * 3 * S);
        ti.o_gnb = (int)o; o += pad16(4LL * 3 * R);
        ti.o_gn = (int)o; o += pad16(8LL * 9 * R);
        ti.o_en = (int)o; o += 16LL * ti.ne_t;
        ti.o_el = (int)o; o += 16LL * ti.ne_t;
        ti.o_egp = (int)o; o += 32LL * ti.ne_t;
        ti.blob_bytes = (int)o;
        if (((long long)p.blob.size() >> 4) + (o >> 4) > 0x7fffffffLL) return "mesh too large for 32-bit blob offsets";
        ti.blob_off = (int)(p.blob.size() >> 4);
        p.blob.resize(p.blob.size() + (size_t)o, 0);
        unsigned char* b = p.blob.data() + ((size_t)ti.blob_off << 4);
        double* cxy = reinterpret_cast<double*>(b + ti.o_cxy);
        for (int l = 0; l < ti.n_l; l++) { const int c = gid(l); cxy[2 * l] = m.cell_cx[c]; cxy[2 * l + 1] = m.cell_cy[c]; }
        double* Sv = reinterpret_cast<double*>(b + ti.o_S);
        for (int l = 0; l < ti.n_g; l++) Sv[l] = m.cell_S[gid(l)];
        for (int l = 0; l < ti.n_l2; l++) b[ti.o_mat + l] = (unsigned char)m.cell_mat[gid(l)];
        uint16_t* slot = reinterpret_cast<uint16_t*>(b + ti.o_slot);
        for (int j = 0; j < n_own; j++)
            for (int k = 0; k < 3; k++) {
                const int e = m.cell_edges[3 * (size_t)(cbeg + j) + k];
                slot[k * S + j] = (uint16_t)(epos[e] * 2 + (m.edge_c2[e] == cbeg + j ? 1 : 0));
            }
        int* gnb = reinterpret_cast<int*>(b + ti.o_gnb);
        double* gn = reinterpret_cast<double*>(b + ti.o_gn);
        for (int r = 0; r < R; r++) {
            const int c = r1c[r];
            for (int k = 0; k < 3; k++) {
                const int e = m.cell_edges[3 * (size_t)c + k];
                const bool is1 = m.edge_c1[e] == c;
                if (!is1 && m.edge_c2[e] != c) return "cell_edges names an edge that does not touch the cell";
                const int nb = other(e, c);
                gnb[k * R + r] = nb >= 0 ? lid(nb) : -1 - m.edge_bc[e];
                gn[(k * 3 + 0) * R + r] = is1 ? m.edge_nx[e] : -m.edge_nx[e];
                gn[(k * 3 + 1) * R + r] = is1 ? m.edge_ny[e] : -m.edge_ny[e];
                gn[(k * 3 + 2) * R + r] = m.edge_l[e];
            }
        }
        double* en = reinterpret_cast<double*>(b + ti.o_en);
        double* egp = reinterpret_cast<double*>(b + ti.o_egp);
        for (int q = 0; q < ti.ne_t; q++) {
            const int e = ekey[q].second;
            const int c1 = m.edge_c1[e], c2 = m.edge_c2[e];
            if (c2 < 0 && !in_tile(c1)) return "internal: boundary edge whose cell is outside the tile";
            if (c2 < 0 && (m.edge_bc[e] < 0 || m.edge_bc[e] > 254)) return "more than 255 boundary conditions";
            const uint32_t l1 = (uint32_t)lid(c1), l2 = c2 >= 0 ? (uint32_t)lid(c2) : (0xff00u | (uint32_t)m.edge_bc[e]);
            if (l1 >= (uint32_t)ti.n_l || (c2 >= 0 && l2 >= (uint32_t)ti.n_l)) return "internal: edge cell outside tile + ring 1";
            en[2 * q] = m.edge_nx[e]; en[2 * q + 1] = m.edge_ny[e];
            unsigned char* el = b + ti.o_el + 16 * (size_t)q;
            const double le = m.edge_l[e];
            const uint32_t cl = l1 | (l2 << 16), zero = 0;
            memcpy(el, &le, 8); memcpy(el + 8, &cl, 4); memcpy(el + 12, &zero, 4);
            for (int i = 0; i < 4; i++) egp[4 * (size_t)q + i] = m.edge_gp[4 * (size_t)e + i];
        }
        p.nl2_max = std::max(p.nl2_max, ti.n_l2); p.nl_max = std::max(p.nl_max, ti.n_l);
        p.ne_max = std::max(p.ne_max, ti.ne_t); p.nring_max = std::max(p.nring_max, ti.n_l2 - n_own);
        p.nhalo_max = std::max(p.nhalo_max, ti.n_l - ti.n_g); p.blob_max = std::max(p.blob_max, ti.blob_bytes);
        p.nrc_max = std::max(p.nrc_max, R);
        p.sum_ring1 += ti.n_l - n_own; p.sum_ring2 += ti.n_l2 - ti.n_l; p.sum_ne += ti.ne_t;
        (ti.halo_dep ? p.boundary : p.interior).push_back(t);
    }
    return "";
}
