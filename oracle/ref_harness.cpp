/* ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * A thin extern "C" window onto the UNMODIFIED reference solver (zhrv/cfd-2d,
 * class FVM_TVD, src/methods/fvm_tvd.{h,cpp}) compiled from the sources where
 * they lie under /root/reference (see oracle/Makefile; output only in
 * oracle/_ref/).  It exists so that
 *   - the C restatement oracle/fvm_oracle.c can be pinned against the real
 *     reference (tests/test_oracle_vs_reference.py, oracle/make_golden.py),
 *   - bench.py has a "kind": "reference" CPU baseline.
 *
 * The harness does NOT change numerics.  The only fix-up is the one SURVEY.md
 * F11 documents: Cell::flag is never initialised by the UNV reader
 * (src/mesh/grid.h:21), so the harness zeroes it right after init().
 *
 * Private/protected members of FVM_TVD are reached with the usual
 * "#define private public" trick; access specifiers do not change layout.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <fstream>
#include <algorithm>
#include <ctime>
#include <unistd.h>
#include <fcntl.h>
#include <typeinfo>

#define private public
#define protected public
#include "fvm_tvd.h"
#undef private
#undef protected

static int g_saved_stdout = -1;
static void quiet_begin(int quiet) {
    if (!quiet) return;
    fflush(stdout);
    g_saved_stdout = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
}
static void quiet_end(int quiet) {
    if (!quiet || g_saved_stdout < 0) return;
    fflush(stdout);
    dup2(g_saved_stdout, 1);
    close(g_saved_stdout);
    g_saved_stdout = -1;
}

static double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

extern "C" {

/* ---- whole-solver access -------------------------------------------------- */

/* chdir(workdir); FVM_TVD::init(xml) (src/methods/fvm_tvd.cpp:7-213); flag=0. */
void* ref_open(const char* workdir, const char* xml, int quiet) {
    if (chdir(workdir) != 0) return NULL;
    if (!hLog) hLog = fopen("task.log", "w");
    Parallel::procCount = 1;
    Parallel::procId = 0;
    FVM_TVD* m = new FVM_TVD();
    quiet_begin(quiet);
    m->init((char*)xml);
    quiet_end(quiet);
    for (int i = 0; i < m->grid.cCount; i++) m->grid.cells[i].flag = 0; /* SURVEY F11 */
    return m;
}

void ref_close(void* h) {
    FVM_TVD* m = (FVM_TVD*)h;
    m->done();
    /* the reference leaks cTau/materials/regions; leave the object itself too */
}

void ref_counts(void* h, int* nc, int* ne, int* nn, int* nmat, int* nbc) {
    FVM_TVD* m = (FVM_TVD*)h;
    *nc = m->grid.cCount; *ne = m->grid.eCount; *nn = m->grid.nCount;
    *nmat = m->matCount; *nbc = m->bCount;
}

/* Flatten Grid (src/mesh/grid.h:18-100) exactly as the product glue does. */
void ref_get_mesh(void* h, double* nodes_xy, int* cell_nodes, int* cell_edges, int* cell_neigh,
                  double* cell_S, double* cell_cx, double* cell_cy, int* cell_mat,
                  int* edge_n1, int* edge_n2, int* edge_c1, int* edge_c2,
                  double* edge_nx, double* edge_ny, double* edge_l, double* edge_gp, int* edge_bc) {
    FVM_TVD* m = (FVM_TVD*)h;
    Grid& g = m->grid;
    for (int i = 0; i < g.nCount; i++) { nodes_xy[2*i] = g.nodes[i].x; nodes_xy[2*i+1] = g.nodes[i].y; }
    for (int i = 0; i < g.cCount; i++) {
        Cell& c = g.cells[i];
        for (int k = 0; k < 3; k++) {
            cell_nodes[3*i+k] = c.nodesInd[k];
            cell_edges[3*i+k] = c.edgesInd[k];
            cell_neigh[3*i+k] = c.neigh[k];
        }
        cell_S[i] = c.S; cell_cx[i] = c.c.x; cell_cy[i] = c.c.y;
        Region& reg = m->getRegion(c.typeName);
        cell_mat[i] = reg.matId;
    }
    for (int e = 0; e < g.eCount; e++) {
        Edge& ed = g.edges[e];
        edge_n1[e] = ed.n1; edge_n2[e] = ed.n2; edge_c1[e] = ed.c1; edge_c2[e] = ed.c2;
        edge_nx[e] = ed.n.x; edge_ny[e] = ed.n.y; edge_l[e] = ed.l;
        edge_gp[4*e+0] = ed.c[1].x; edge_gp[4*e+1] = ed.c[1].y;
        edge_gp[4*e+2] = ed.c[2].x; edge_gp[4*e+3] = ed.c[2].y;
        int ib = -1;
        if (ed.bnd) for (int b = 0; b < m->bCount; b++) if (m->boundaries[b] == ed.bnd) ib = b;
        edge_bc[e] = ib;
    }
}

/* kind: 1 inlet, 2 outlet, 3 wall (slip and "no-slip" are the same class,
 * src/bnd_cond.cpp:50-62). */
void ref_get_phys(void* h, double* mat_M, double* mat_Cp, int* bc_kind, double* bc_par,
                  double* limits5, double* cfl, double* tau, int* steady) {
    FVM_TVD* m = (FVM_TVD*)h;
    for (int i = 0; i < m->matCount; i++) { mat_M[i] = m->materials[i].M; mat_Cp[i] = m->materials[i].Cp; }
    for (int b = 0; b < m->bCount; b++) {
        CFDBoundary* bc = m->boundaries[b];
        int kind = 0;
        if (dynamic_cast<CFDBndInlet*>(bc)) kind = 1;
        else if (dynamic_cast<CFDBndOutlet*>(bc)) kind = 2;
        else if (dynamic_cast<CFDBndWallSlip*>(bc)) kind = 3;
        else if (dynamic_cast<CFDBndWallNoSlip*>(bc)) kind = 3;
        bc_kind[b] = kind;
        for (int k = 0; k < 4; k++) bc_par[4*b+k] = (bc->par && k < bc->parCount) ? bc->par[k] : 0.0;
    }
    limits5[0] = m->limitRmin; limits5[1] = m->limitRmax; limits5[2] = m->limitPmin;
    limits5[3] = m->limitPmax; limits5[4] = m->limitUmax;
    *cfl = m->CFL; *tau = m->TAU; *steady = m->STEADY ? 1 : 0;
}

void ref_set_state(void* h, const double* ro, const double* ru, const double* rv, const double* re) {
    FVM_TVD* m = (FVM_TVD*)h;
    size_t n = m->grid.cCount * sizeof(double);
    memcpy(m->ro, ro, n); memcpy(m->ru, ru, n); memcpy(m->rv, rv, n); memcpy(m->re, re, n);
    memcpy(m->ro_old, ro, n); memcpy(m->ru_old, ru, n); memcpy(m->rv_old, rv, n); memcpy(m->re_old, re, n);
}

void ref_set_flags(void* h, const unsigned int* flag) {
    FVM_TVD* m = (FVM_TVD*)h;
    for (int i = 0; i < m->grid.cCount; i++) m->grid.cells[i].flag = flag[i];
}

void ref_set_control(void* h, double tau, double cfl, int steady) {
    FVM_TVD* m = (FVM_TVD*)h;
    m->TAU = tau; m->CFL = cfl; m->STEADY = steady != 0;
}

void ref_set_limits(void* h, const double* limits5) {
    FVM_TVD* m = (FVM_TVD*)h;
    m->limitRmin = limits5[0]; m->limitRmax = limits5[1]; m->limitPmin = limits5[2];
    m->limitPmax = limits5[3]; m->limitUmax = limits5[4];
}

/* FVM_TVD::calcTimeStep (src/methods/fvm_tvd.cpp:216-240). Returns TAU. */
double ref_calc_time_step(void* h, int quiet) {
    FVM_TVD* m = (FVM_TVD*)h;
    quiet_begin(quiet);
    m->calcTimeStep();
    quiet_end(quiet);
    return m->TAU;
}

/* FVM_TVD::run (src/methods/fvm_tvd.cpp:303-462) for exactly nsteps steps with
 * file and log output switched off.  Returns wall seconds of run() alone. */
double ref_run(void* h, int nsteps, int quiet) {
    FVM_TVD* m = (FVM_TVD*)h;
    m->STEP_MAX = nsteps;
    m->TMAX = 1.0e300;
    m->FILE_SAVE_STEP = 0x7fffffff;
    m->PRINT_STEP = 0x7fffffff;
    quiet_begin(quiet);
    double t0 = now_s();
    m->run();
    double t1 = now_s();
    quiet_end(quiet);
    return t1 - t0;
}

void ref_get_state(void* h, double* ro, double* ru, double* rv, double* re, double* cTau, unsigned int* flag) {
    FVM_TVD* m = (FVM_TVD*)h;
    size_t n = m->grid.cCount * sizeof(double);
    memcpy(ro, m->ro, n); memcpy(ru, m->ru, n); memcpy(rv, m->rv, n); memcpy(re, m->re, n);
    if (cTau) memcpy(cTau, m->cTau, n);
    if (flag) for (int i = 0; i < m->grid.cCount; i++) flag[i] = m->grid.cells[i].flag;
}

/* FVM_TVD::calcGrad (src/methods/fvm_tvd.cpp:242-301); grads as [nc][8] =
 * (Rx,Ry,Px,Py,Ux,Uy,Vx,Vy). */
void ref_calc_grad(void* h, double* grad8) {
    FVM_TVD* m = (FVM_TVD*)h;
    m->calcGrad();
    for (int i = 0; i < m->grid.cCount; i++) {
        grad8[8*i+0] = m->gradR[i].x; grad8[8*i+1] = m->gradR[i].y;
        grad8[8*i+2] = m->gradP[i].x; grad8[8*i+3] = m->gradP[i].y;
        grad8[8*i+4] = m->gradU[i].x; grad8[8*i+5] = m->gradU[i].y;
        grad8[8*i+6] = m->gradV[i].x; grad8[8*i+7] = m->gradV[i].y;
    }
}

/* One stage's edge fluxes exactly as run() forms them (fvm_tvd.cpp:329-352):
 * calcGrad must have been called.  flux4 = [ne][4] (fr,fu,fv,fe summed over GPs). */
void ref_edge_fluxes(void* h, double* flux4) {
    FVM_TVD* m = (FVM_TVD*)h;
    Grid& grid = m->grid;
    for (int iEdge = 0; iEdge < grid.eCount; iEdge++) {
        Vector n = grid.edges[iEdge].n;
        Param pL, pR;
        double fr = 0.0, fu = 0.0, fv = 0.0, fe = 0.0;
        for (int iGP = 1; iGP < grid.edges[iEdge].cCount; iGP++) {
            double fr1, fu1, fv1, fe1;
            m->reconstruct(iEdge, pL, pR, grid.edges[iEdge].c[iGP]);
            m->calcFlux(fr1, fu1, fv1, fe1, pL, pR, n, 1.4);
            fr += fr1; fu += fu1; fv += fv1; fe += fe1;
        }
        flux4[4*iEdge+0] = fr; flux4[4*iEdge+1] = fu; flux4[4*iEdge+2] = fv; flux4[4*iEdge+3] = fe;
    }
}

/* FVM_TVD::convertConsToPar for one cell; out = r,p,e,E,u,v,cz,T. */
void ref_cons_to_par(void* h, int iCell, double* out8) {
    FVM_TVD* m = (FVM_TVD*)h;
    Param p; memset(&p, 0, sizeof p);
    m->convertConsToPar(iCell, p);
    out8[0] = p.r; out8[1] = p.p; out8[2] = p.e; out8[3] = p.E; out8[4] = p.u; out8[5] = p.v; out8[6] = p.cz; out8[7] = p.T;
}

/* FVM_TVD::boundaryCond (fvm_tvd.cpp:694-711) on a boundary edge; pL8/pR8 as above. */
void ref_boundary_cond(void* h, int iEdge, const double* pL8, double* pR8) {
    FVM_TVD* m = (FVM_TVD*)h;
    Param L, R; memset(&R, 0, sizeof R); memset(&L, 0, sizeof L);
    L.r = pL8[0]; L.p = pL8[1]; L.e = pL8[2]; L.E = pL8[3]; L.u = pL8[4]; L.v = pL8[5]; L.cz = pL8[6]; L.T = pL8[7];
    m->boundaryCond(iEdge, L, R);
    pR8[0] = R.r; pR8[1] = R.p; pR8[2] = R.e; pR8[3] = R.E; pR8[4] = R.u; pR8[5] = R.v; pR8[6] = R.cz; pR8[7] = R.T;
}

/* ---- function-level known-answer access ----------------------------------- */

/* rim_orig (src/global.cpp:232-405), WB=WE=0 as FVM_TVD::calcFlux calls it.
 * in: [n][8] = RB,PB,UB,VB,RE,PE,UE,VE ; out: [n][5] = RI,EI,PI,UI,VI */
void ref_rim_orig(int n, const double* in8, double gam, double* out5) {
    for (int i = 0; i < n; i++) {
        const double* a = in8 + 8*i;
        double RI, EI, PI, UI, VI, WI;
        rim_orig(RI, EI, PI, UI, VI, WI, a[0], a[1], a[2], a[3], 0, a[4], a[5], a[6], a[7], 0, gam);
        out5[5*i+0] = RI; out5[5*i+1] = EI; out5[5*i+2] = PI; out5[5*i+3] = UI; out5[5*i+4] = VI;
    }
}

/* FVM_TVD::calcFlux (fvm_tvd.cpp:602-643) as compiled into this variant.
 * in: [n][12] = rL,pL,uL,vL,EL, rR,pR,uR,vR,ER, nx,ny ; out [n][4]. */
void ref_calc_flux(int n, const double* in12, double gam, double* out4) {
    FVM_TVD m;
    for (int i = 0; i < n; i++) {
        const double* a = in12 + 12*i;
        Param L, R; memset(&L, 0, sizeof L); memset(&R, 0, sizeof R);
        L.r = a[0]; L.p = a[1]; L.u = a[2]; L.v = a[3]; L.E = a[4];
        R.r = a[5]; R.p = a[6]; R.u = a[7]; R.v = a[8]; R.E = a[9];
        Vector nn; nn.x = a[10]; nn.y = a[11];
        m.calcFlux(out4[4*i], out4[4*i+1], out4[4*i+2], out4[4*i+3], L, R, nn, gam);
    }
}

/* Material::URS (src/global.cpp:9-30). io: [n][8] = r,p,e,E,u,v,cz,T (in place). */
void ref_urs(int n, double M, double Cp, int mode, double* io8) {
    Material mat; mat.M = M; mat.Cp = Cp;
    for (int i = 0; i < n; i++) {
        double* a = io8 + 8*i;
        Param p; memset(&p, 0, sizeof p);
        p.r = a[0]; p.p = a[1]; p.e = a[2]; p.E = a[3]; p.u = a[4]; p.v = a[5]; p.cz = a[6]; p.T = a[7];
        mat.URS(p, mode);
        a[0] = p.r; a[1] = p.p; a[2] = p.e; a[3] = p.E; a[4] = p.u; a[5] = p.v; a[6] = p.cz; a[7] = p.T;
    }
}

const char* ref_variant(void) {
#ifdef REF_VARIANT
    return REF_VARIANT;
#else
    return "v0";
#endif
}

} /* extern "C" */
