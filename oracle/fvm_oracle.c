/* fvm_oracle.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * CPU restatement, in plain C99, of the explicit finite-volume path of zhrv/cfd-2d
 * (class FVM_TVD).  It is the checker the CUDA path is compared with on the GPU box, where
 * /root/reference does not exist.  PARITY IS PINNED: tests/test_oracle_golden.py (against the committed golden vectors) and
 * oracle/make_golden.py compare every function below BIT FOR BIT with the real reference
 * compiled from /root/reference (oracle/_ref, oracle/ref_harness.cpp); the golden vectors those
 * runs produce are committed under tests/golden/.  (The reference itself ships no tests or
 * golden vectors for this path -- SURVEY.md section 4.)
 *
 * Build: gcc -std=c99 -O2 -ffp-contract=off  (no FMA contraction: the reference's x86-64 -O2
 * build has none).
 *
 * Loop structure follows the reference (edge-ordered scatter loops), NOT the GPU's gather
 * formulation, so the two are independent statements of the same arithmetic.
 *
 * Reference map
 *   prim()            FVM_TVD::convertConsToPar  src/methods/fvm_tvd.cpp:803-813
 *                     Material::URS modes 0,1,2  src/global.cpp:9-30
 *   ghost()           FVM_TVD::boundaryCond      fvm_tvd.cpp:694-711 ; CFDBnd*::run src/bnd_cond.cpp:75-110
 *   calc_grad()       FVM_TVD::calcGrad          fvm_tvd.cpp:242-301
 *   reconstruct()     FVM_TVD::reconstruct       fvm_tvd.cpp:646-691
 *   riemann()         rim_orig                   src/global.cpp:232-405
 *   flux_godunov()    FVM_TVD::calcFlux          fvm_tvd.cpp:604-622
 *   flux_lax()        the commented LF block     fvm_tvd.cpp:623-642
 *   stage()           RK sub-step of run()       fvm_tvd.cpp:323-374 (= :376-427)
 *   fvm_oracle_step() FVM_TVD::run loop body     fvm_tvd.cpp:310-450
 *   remediate()       remediateLimCells          fvm_tvd.cpp:464-499
 *   calc_time_step()  FVM_TVD::calcTimeStep      fvm_tvd.cpp:216-240
 *
 * Deliberate differences from the reference (documented in DESIGN.md):
 *   - Cell::flag starts at 0 (the reference leaves it uninitialised);
 *   - the Newton loop of rim_orig is capped (the reference loops forever on bad states, F3).
 */
#include "fvm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GR 8.314472 /* Material::gR, global.cpp:6 */

typedef struct { double r, p, e, E, u, v, cz, T; } Param; /* global.h:180-194 (ML unused) */

struct fvm_oracle {
    int nc, ne, nmat, nbc;
    double *S, *cx, *cy; int *mat; int *cedges;
    int *c1, *c2, *bc; double *nx, *ny, *l, *gp;
    double *mM, *mCp; int *bkind; double *bpar; double lim[5];
    double CFL, TAU, t; int steady, flux, order, max_newton;
    double *ro, *ru, *rv, *re, *ro_old, *ru_old, *rv_old, *re_old, *ro_int, *ru_int, *rv_int, *re_int;
    double *cTau; double *grad; /* [nc][8] Rx,Ry,Px,Py,Ux,Uy,Vx,Vy */
    uint32_t* flag;
    long long newton_iters, riemann_calls; int err;
};

static void* dup_arr(const void* src, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (src && bytes) memcpy(p, src, bytes);
    return p;
}

fvm_oracle* fvm_oracle_create(const cfd2d_mesh* m, const cfd2d_phys* p, const cfd2d_ctrl* c) {
    fvm_oracle* o = (fvm_oracle*)calloc(1, sizeof *o);
    int nc = m->nc, ne = m->ne;
    o->nc = nc; o->ne = ne; o->nmat = p->nmat; o->nbc = p->nbc;
    o->S = dup_arr(m->cell_S, nc * 8); o->cx = dup_arr(m->cell_cx, nc * 8); o->cy = dup_arr(m->cell_cy, nc * 8);
    o->mat = dup_arr(m->cell_mat, nc * 4); o->cedges = dup_arr(m->cell_edges, nc * 12);
    o->c1 = dup_arr(m->edge_c1, ne * 4); o->c2 = dup_arr(m->edge_c2, ne * 4); o->bc = dup_arr(m->edge_bc, ne * 4);
    o->nx = dup_arr(m->edge_nx, ne * 8); o->ny = dup_arr(m->edge_ny, ne * 8); o->l = dup_arr(m->edge_l, ne * 8);
    o->gp = dup_arr(m->edge_gp, ne * 32);
    o->mM = dup_arr(p->mat_M, p->nmat * 8); o->mCp = dup_arr(p->mat_Cp, p->nmat * 8);
    o->bkind = dup_arr(p->bc_kind, p->nbc * 4); o->bpar = dup_arr(p->bc_par, p->nbc * 32);
    memcpy(o->lim, p->limits, sizeof o->lim);
    o->CFL = c->CFL; o->TAU = c->TAU; o->steady = c->steady; o->flux = c->flux; o->order = c->order;
    o->max_newton = c->max_newton > 0 ? c->max_newton : 1000;
    double** st[] = { &o->ro, &o->ru, &o->rv, &o->re, &o->ro_old, &o->ru_old, &o->rv_old, &o->re_old,
                      &o->ro_int, &o->ru_int, &o->rv_int, &o->re_int, &o->cTau };
    for (unsigned i = 0; i < sizeof st / sizeof st[0]; i++) *st[i] = (double*)calloc(nc ? nc : 1, 8);
    o->grad = (double*)calloc(nc ? nc : 1, 64);
    o->flag = (uint32_t*)calloc(nc ? nc : 1, 4);
    return o;
}

void fvm_oracle_destroy(fvm_oracle* o) {
    if (!o) return;
    void* a[] = { o->S, o->cx, o->cy, o->mat, o->cedges, o->c1, o->c2, o->bc, o->nx, o->ny, o->l, o->gp, o->mM, o->mCp,
                  o->bkind, o->bpar, o->ro, o->ru, o->rv, o->re, o->ro_old, o->ru_old, o->rv_old, o->re_old,
                  o->ro_int, o->ru_int, o->rv_int, o->re_int, o->cTau, o->grad, o->flag };
    for (unsigned i = 0; i < sizeof a / sizeof a[0]; i++) free(a[i]);
    free(o);
}

void fvm_oracle_set_state(fvm_oracle* o, const double* ro, const double* ru, const double* rv,
                          const double* re, const uint32_t* flag) {
    size_t n = (size_t)o->nc * 8;
    memcpy(o->ro, ro, n); memcpy(o->ru, ru, n); memcpy(o->rv, rv, n); memcpy(o->re, re, n);
    memcpy(o->ro_old, ro, n); memcpy(o->ru_old, ru, n); memcpy(o->rv_old, rv, n); memcpy(o->re_old, re, n);
    if (flag) memcpy(o->flag, flag, (size_t)o->nc * 4); else memset(o->flag, 0, (size_t)o->nc * 4);
}

void fvm_oracle_get_state(fvm_oracle* o, double* ro, double* ru, double* rv, double* re, double* cTau, uint32_t* flag) {
    size_t n = (size_t)o->nc * 8;
    memcpy(ro, o->ro, n); memcpy(ru, o->ru, n); memcpy(rv, o->rv, n); memcpy(re, o->re, n);
    if (cTau) memcpy(cTau, o->cTau, n);
    if (flag) memcpy(flag, o->flag, (size_t)o->nc * 4);
}

/* Material::URS, global.cpp:9-30 */
static void urs(const fvm_oracle* o, int imat, Param* par, int mode) {
    double M = o->mM[imat], Cp = o->mCp[imat];
    double Cv = Cp - GR / M;
    double gam = Cp / Cv;
    switch (mode) {
    case 0: par->p = par->r * par->e * (gam - 1); par->cz = sqrt(gam * par->p / par->r); break;
    case 1: par->e = par->p / (par->r * (gam - 1)); par->T = par->e / Cv; break;
    case 2: par->r = par->p * M / (par->T * GR); par->cz = sqrt(gam * par->p / par->r); break;
    }
}

/* known-answer hook: Material::URS on io8[n][8] = r,p,e,E,u,v,cz,T in place */
void fvm_oracle_urs(int n, double M, double Cp, int mode, double* io8) {
    fvm_oracle o;
    memset(&o, 0, sizeof o);
    o.mM = &M; o.mCp = &Cp;
    for (int i = 0; i < n; i++) {
        double* a = io8 + 8 * (size_t)i;
        Param p = { a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7] };
        urs(&o, 0, &p, mode);
        a[0] = p.r; a[1] = p.p; a[2] = p.e; a[3] = p.E; a[4] = p.u; a[5] = p.v; a[6] = p.cz; a[7] = p.T;
    }
}

/* FVM_TVD::convertConsToPar, fvm_tvd.cpp:803-813 */
static void prim(const fvm_oracle* o, int c, Param* par) {
    par->r = o->ro[c];
    par->u = o->ru[c] / o->ro[c];
    par->v = o->rv[c] / o->ro[c];
    par->E = o->re[c] / o->ro[c];
    par->e = par->E - 0.5 * (par->u * par->u + par->v * par->v);
    urs(o, o->mat[c], par, 0);
    urs(o, o->mat[c], par, 1);
}

/* FVM_TVD::boundaryCond fvm_tvd.cpp:694-711 + CFDBnd*::run bnd_cond.cpp:75-110 */
static void ghost(const fvm_oracle* o, int e, const Param* pL, Param* pR) {
    int ib = o->bc[e];
    int kind = o->bkind[ib];
    const double* par = o->bpar + 4 * ib;
    if (kind == CFD2D_BC_INLET) {
        /* the reference leaves the other members of pR as they were; all of them are
         * overwritten below before anything reads them */
        pR->u = par[0]; pR->v = par[1]; pR->T = par[2]; pR->p = par[3];
    } else if (kind == CFD2D_BC_OUTLET) {
        *pR = *pL;
    } else { /* wall: slip and "no-slip" are the same class */
        *pR = *pL;
        double Un = pL->u * o->nx[e] + pL->v * o->ny[e];
        double Vx = o->nx[e] * Un * 2.0;
        double Vy = o->ny[e] * Un * 2.0;
        pR->u = pL->u - Vx;
        pR->v = pL->v - Vy;
    }
    int im = o->mat[o->c1[e]];
    urs(o, im, pR, 2);
    urs(o, im, pR, 1);
    pR->E = pR->e + 0.5 * (pR->u * pR->u + pR->v * pR->v);
}

/* FVM_TVD::calcGrad, fvm_tvd.cpp:242-301 */
static void calc_grad(fvm_oracle* o) {
    double* g = o->grad;
    memset(g, 0, (size_t)o->nc * 64);
    for (int e = 0; e < o->ne; e++) {
        int c1 = o->c1[e], c2 = o->c2[e];
        Param pL, pR;
        memset(&pR, 0, sizeof pR);
        prim(o, c1, &pL);
        if (c2 > -1) prim(o, c2, &pR); else ghost(o, e, &pL, &pR);
        double nx = o->nx[e], ny = o->ny[e], l = o->l[e];
        double* a = g + 8 * (size_t)c1;
        a[0] += (pL.r + pR.r) / 2 * nx * l;  a[1] += (pL.r + pR.r) / 2 * ny * l;
        a[2] += (pL.p + pR.p) / 2 * nx * l;  a[3] += (pL.p + pR.p) / 2 * ny * l;
        a[4] += (pL.u + pR.u) / 2 * nx * l;  a[5] += (pL.u + pR.u) / 2 * ny * l;
        a[6] += (pL.v + pR.v) / 2 * nx * l;  a[7] += (pL.v + pR.v) / 2 * ny * l;
        if (c2 > -1) {
            double* b = g + 8 * (size_t)c2;
            b[0] -= (pL.r + pR.r) / 2 * nx * l;  b[1] -= (pL.r + pR.r) / 2 * ny * l;
            b[2] -= (pL.p + pR.p) / 2 * nx * l;  b[3] -= (pL.p + pR.p) / 2 * ny * l;
            b[4] -= (pL.u + pR.u) / 2 * nx * l;  b[5] -= (pL.u + pR.u) / 2 * ny * l;
            b[6] -= (pL.v + pR.v) / 2 * nx * l;  b[7] -= (pL.v + pR.v) / 2 * ny * l;
        }
    }
    for (int c = 0; c < o->nc; c++) {
        double si = o->S[c];
        for (int k = 0; k < 8; k++) g[8 * (size_t)c + k] /= si;
    }
}

/* FVM_TVD::reconstruct, fvm_tvd.cpp:646-691.  order==1 is the "//return;" variant (:654, :678). */
static void reconstruct(const fvm_oracle* o, int e, Param* pL, Param* pR, double PEx, double PEy) {
    int c1 = o->c1[e], c2 = o->c2[e];
    const double* g1 = o->grad + 8 * (size_t)c1;
    if (c2 > -1) { /* Edge::TYPE_INNER */
        prim(o, c1, pL);
        prim(o, c2, pR);
        if (o->order == 1) return;
        const double* g2 = o->grad + 8 * (size_t)c2;
        double DL1x = PEx - o->cx[c1], DL1y = PEy - o->cy[c1];
        double DL2x = PEx - o->cx[c2], DL2y = PEy - o->cy[c2];
        pL->r += g1[0] * DL1x + g1[1] * DL1y;
        pL->p += g1[2] * DL1x + g1[3] * DL1y;
        pL->u += g1[4] * DL1x + g1[5] * DL1y;
        pL->v += g1[6] * DL1x + g1[7] * DL1y;
        pR->r += g2[0] * DL2x + g2[1] * DL2y;
        pR->p += g2[2] * DL2x + g2[3] * DL2y;
        pR->u += g2[4] * DL2x + g2[5] * DL2y;
        pR->v += g2[6] * DL2x + g2[7] * DL2y;
    } else {
        prim(o, c1, pL);
        if (o->order != 1) {
            double DL1x = PEx - o->cx[c1], DL1y = PEy - o->cy[c1];
            pL->r += g1[0] * DL1x + g1[1] * DL1y;
            pL->p += g1[2] * DL1x + g1[3] * DL1y;
            pL->u += g1[4] * DL1x + g1[5] * DL1y;
            pL->v += g1[6] * DL1x + g1[7] * DL1y;
        }
        ghost(o, e, pL, pR);
    }
}

/* rim_orig, global.cpp:232-405 (WB = WE = 0; WI dropped).  Returns Newton iterations, or -1 if
 * the cap was hit. */
static int riemann(double* RI, double* EI, double* PI, double* UI, double* VI,
                   double RB, double PB, double UB, double VB,
                   double RE, double PE, double UE, double VE, double GAM, int max_newton) {
    double AGAM = (GAM - 1.0);
    double DGAM = (2.0 / AGAM);
    double GGAM = (sqrt(GAM * AGAM));
    double HGAM = (AGAM / 2.0);
    double FGAM = (3.0 * GAM - 1.0);
    double OGAM = (AGAM / (2.0 * GAM));
    double QGAM = (GAM + 1.0);
    double PGAM = (QGAM / (2.0 * GAM));
    double RGAM = (4.0 * GAM);
    double SGAM = (GAM * AGAM);
    double TGAM = (QGAM / 2.0);
    double US = 0.0, UF = 0.0;
    double RF = 0, RS = 0, EF = 0, ES = 0, SBL, SFL, SSL, SEL, D;
    double PPB, PKB, ZNB, F1, FS1, ZFB, PPE, PKE, ZNE, F2, FS2, ZFE, DP, UBD, RUBD, UED, RUED, P;
    double eps = 1.0e-5;
    double CB = sqrt(GAM * PB / RB);
    double CE = sqrt(GAM * PE / RE);
    double EB = CB * CB / SGAM;
    double EE = CE * CE / SGAM;
    double RCB = RB * CB;
    double RCE = RE * CE;
    double DU = UB - UE;
    int it = 0;
    if (DU < -2.0 * (CB + CE) / AGAM) { /* vacuum */
        RF = 0.0; RS = 0.0; EF = 0.0; ES = 0.0;
        SBL = UB - CB;
        SFL = UB + 2.0 * CB / AGAM;
        SSL = UE - 2.0 * CE / AGAM;
        SEL = UE + CE;
    } else {
        P = (PB * RCE + PE * RCB + DU * RCB * RCE) / (RCB + RCE);
        for (;;) {
            if (P < eps) P = eps;
            PPB = P / PB;
            if (PB > P) {
                ZFB = CB * exp(log(PPB) * OGAM);
                F1 = DGAM * (ZFB - CB);
                FS1 = ZFB / (GAM * P);
            } else {
                PKB = PGAM * PPB + OGAM;
                ZNB = RCB * sqrt(PKB);
                F1 = (P - PB) / ZNB;
                FS1 = (QGAM * PPB + FGAM) / (RGAM * ZNB * PKB);
            }
            PPE = P / PE;
            if (PE > P) {
                ZFE = CE * exp(log(PPE) * OGAM);
                F2 = DGAM * (ZFE - CE);
                FS2 = ZFE / (GAM * P);
            } else {
                PKE = PGAM * PPE + OGAM;
                ZNE = RCE * sqrt(PKE);
                F2 = (P - PE) / ZNE;
                FS2 = (QGAM * PPE + FGAM) / (RGAM * ZNE * PKE);
            }
            DP = (DU - F1 - F2) / (FS1 + FS2);
            P = P + DP;
            it++;
            if (!(fabs(DU - F1 - F2) > eps)) break;
            if (it >= max_newton) { it = -1; break; }
        }
        PPB = P / PB;
        PPE = P / PE;
        ZFB = CB * exp(log(PPB) * OGAM);
        ZFE = CE * exp(log(PPE) * OGAM);
        if (PB > P) {
            EF = ZFB * ZFB / SGAM;
            UF = UB + DGAM * (CB - ZFB);
            RF = P / (AGAM * EF);
            SBL = UB - CB;
            SFL = UF - ZFB;
        } else {
            D = UB - sqrt((TGAM * P + HGAM * PB) / RB);
            UBD = UB - D;
            RUBD = RB * UBD;
            RF = RUBD * RUBD / (PB - P + RUBD * UBD);
            UF = D + RUBD / RF;
            EF = P / (AGAM * RF);
            SBL = D;
            SFL = D;
        }
        if (PE > P) {
            ES = ZFE * ZFE / SGAM;
            US = UE - DGAM * (CE - ZFE);
            RS = P / (AGAM * ES);
            SSL = US + ZFE;
            SEL = UE + CE;
        } else {
            D = UE + sqrt((TGAM * P + HGAM * PE) / RE);
            UED = UE - D;
            RUED = RE * UED;
            RS = RUED * RUED / (PE - P + RUED * UED);
            US = D + RUED / RS;
            ES = P / (AGAM * RS);
            SEL = D;
            SSL = D;
        }
    }
    /* sampling at x/t = 0 */
    if (SEL <= 0.0) {
        *RI = RE; *EI = EE; *UI = UE; *VI = VE;
    } else if (SBL >= 0.0) {
        *RI = RB; *EI = EB; *UI = UB; *VI = VB;
    } else if ((SSL >= 0.0) && (SFL <= 0.0)) {
        if (US >= 0.0) { *RI = RF; *EI = EF; *UI = UF; *VI = VB; }
        else           { *RI = RS; *EI = ES; *UI = US; *VI = VE; }
    } else if (SFL > 0.0) {
        *UI = (UB + DGAM * GGAM * sqrt(EB)) / (1 + DGAM);
        *VI = VB;
        *EI = ((*UI) * (*UI)) / SGAM;
        *RI = RB * exp(log(*EI / EB) * (1 / AGAM));
    } else {
        *UI = (UE - DGAM * GGAM * sqrt(EE)) / (1 + DGAM);
        *VI = VE;
        *EI = ((*UI) * (*UI)) / SGAM;
        *RI = RE * exp(log(*EI / EE) * (1 / AGAM));
    }
    *PI = AGAM * (*EI) * (*RI);
    return it;
}

/* FVM_TVD::calcFlux Godunov block, fvm_tvd.cpp:604-622 */
static int flux_godunov(double* fr, double* fu, double* fv, double* fe, const Param* pL, const Param* pR,
                        double nx, double ny, double GAM, int max_newton) {
    double RI, EI, PI, UI, VI, UN, UT;
    double unl = pL->u * nx + pL->v * ny;
    double unr = pR->u * nx + pR->v * ny;
    double utl = pL->u * ny - pL->v * nx;
    double utr = pR->u * ny - pR->v * nx;
    int it = riemann(&RI, &EI, &PI, &UN, &UT, pL->r, pL->p, unl, utl, pR->r, pR->p, unr, utr, GAM, max_newton);
    UI = UN * nx + UT * ny;
    VI = UN * ny - UT * nx;
    *fr = RI * UN;
    *fu = *fr * UI + PI * nx;
    *fv = *fr * VI + PI * ny;
    *fe = (RI * (EI + 0.5 * (UI * UI + VI * VI)) + PI) * UN;
    return it;
}

static double max2(double a, double b) { if (a > b) return a; else return b; } /* _max_, grid.h:102 */

/* the commented Lax-Friedrichs block of FVM_TVD::calcFlux, fvm_tvd.cpp:623-642 */
static void flux_lax(double* fr, double* fu, double* fv, double* fe, const Param* pL, const Param* pR,
                     double nx, double ny, double GAM) {
    double unl = pL->u * nx + pL->v * ny;
    double unr = pR->u * nx + pR->v * ny;
    double rol, rul, rvl, rel, ror, rur, rvr, rer;
    double alpha = max2(fabs(unl) + sqrt(GAM * pL->p / pL->r), fabs(unr) + sqrt(GAM * pR->p / pR->r));
    rol = pL->r; rul = pL->r * pL->u; rvl = pL->r * pL->v; rel = pL->r * pL->E;
    ror = pR->r; rur = pR->r * pR->u; rvr = pR->r * pR->v; rer = pR->r * pR->E;
    double frl = rol * unl;
    double frr = ror * unr;
    *fr = 0.5 * (frr + frl - alpha * (ror - rol));
    *fu = 0.5 * (frr * pR->u + frl * pL->u + (pR->p + pL->p) * nx - alpha * (rur - rul));
    *fv = 0.5 * (frr * pR->v + frl * pL->v + (pR->p + pL->p) * ny - alpha * (rvr - rvl));
    *fe = 0.5 * ((rer + pR->p) * unr + (rel + pL->p) * unl - alpha * (rer - rel));
}

/* flux of one edge summed over its two Gauss points, fvm_tvd.cpp:331-352 */
static void edge_flux(fvm_oracle* o, int e, double F[4]) {
    double fr = 0.0, fu = 0.0, fv = 0.0, fe = 0.0;
    Param pL, pR;
    memset(&pR, 0, sizeof pR);
    for (int iGP = 1; iGP < 3; iGP++) {
        double fr1, fu1, fv1, fe1;
        reconstruct(o, e, &pL, &pR, o->gp[4 * (size_t)e + 2 * (iGP - 1)], o->gp[4 * (size_t)e + 2 * (iGP - 1) + 1]);
        double GAM = 1.4; /* fvm_tvd.cpp:345 */
        if (o->flux == CFD2D_FLUX_LAX) {
            flux_lax(&fr1, &fu1, &fv1, &fe1, &pL, &pR, o->nx[e], o->ny[e], GAM);
        } else {
            int it = flux_godunov(&fr1, &fu1, &fv1, &fe1, &pL, &pR, o->nx[e], o->ny[e], GAM, o->max_newton);
            o->riemann_calls++;
            if (it < 0) { o->err = CFD2D_ENEWTON; o->newton_iters += o->max_newton; } else o->newton_iters += it;
        }
        fr += fr1; fu += fu1; fv += fv1; fe += fe1;
    }
    F[0] = fr; F[1] = fu; F[2] = fv; F[3] = fe;
}

/* one RK sub-step, fvm_tvd.cpp:323-374 */
static void stage(fvm_oracle* o) {
    size_t n = (size_t)o->nc * 8;
    memset(o->ro_int, 0, n); memset(o->ru_int, 0, n); memset(o->rv_int, 0, n); memset(o->re_int, 0, n);
    calc_grad(o);
    for (int e = 0; e < o->ne; e++) {
        int c1 = o->c1[e], c2 = o->c2[e];
        double l = o->l[e] * 0.5;
        double F[4];
        edge_flux(o, e, F);
        o->ro_int[c1] -= F[0] * l; o->ru_int[c1] -= F[1] * l; o->rv_int[c1] -= F[2] * l; o->re_int[c1] -= F[3] * l;
        if (c2 > -1) {
            o->ro_int[c2] += F[0] * l; o->ru_int[c2] += F[1] * l; o->rv_int[c2] += F[2] * l; o->re_int[c2] += F[3] * l;
        }
    }
    for (int c = 0; c < o->nc; c++) {
        if ((o->flag[c] & CFD2D_CELL_FLAG_LIM) > 0) continue;
        double cfl = o->cTau[c] / o->S[c];
        o->ro[c] += cfl * o->ro_int[c];
        o->ru[c] += cfl * o->ru_int[c];
        o->rv[c] += cfl * o->rv_int[c];
        o->re[c] += cfl * o->re_int[c];
    }
}

/* FVM_TVD::remediateLimCells, fvm_tvd.cpp:464-499 (including the c2-only neighbour quirk) */
static void remediate(fvm_oracle* o) {
    for (int c = 0; c < o->nc; c++) {
        if ((o->flag[c] & CFD2D_CELL_FLAG_LIM) > 0) {
            double sRO = 0.0, sRU = 0.0, sRV = 0.0, sRE = 0.0, S = 0.0;
            for (int i = 0; i < 3; i++) {
                int e = o->cedges[3 * (size_t)c + i];
                int j = o->c2[e];
                if (j >= 0) {
                    double s = o->S[j];
                    S += s;
                    sRO += o->ro[j] * s; sRU += o->ru[j] * s; sRV += o->rv[j] * s; sRE += o->re[j] * s;
                }
            }
            o->ro[c] = sRO / S; o->ru[c] = sRU / S; o->rv[c] = sRV / S; o->re[c] = sRE / S;
            o->flag[c] += 0x010000;
            if (o->flag[c] & 0x200000) o->flag[c] &= 0x001110;
        }
    }
}

static double cell_tau(const fvm_oracle* o, int c) {
    Param p;
    prim(o, c, &p);
    return o->CFL * o->S[c] / max2(fabs(p.u) + p.cz, fabs(p.v) + p.cz);
}

/* FVM_TVD::calcTimeStep, fvm_tvd.cpp:216-240 */
double fvm_oracle_calc_time_step(fvm_oracle* o) {
    if (o->steady) {
        for (int c = 0; c < o->nc; c++) o->cTau[c] = cell_tau(o, c);
    } else {
        for (int c = 0; c < o->nc; c++) {
            double t = cell_tau(o, c);
            if (o->TAU > t) o->TAU = t;
        }
        for (int c = 0; c < o->nc; c++) o->cTau[c] = o->TAU;
    }
    return o->TAU;
}

/* body of the while loop of FVM_TVD::run, fvm_tvd.cpp:310-450 */
int fvm_oracle_step(fvm_oracle* o, int nsteps) {
    size_t n = (size_t)o->nc * 8;
    for (int s = 0; s < nsteps; s++) {
        if (!o->steady) o->t += o->TAU; else fvm_oracle_calc_time_step(o);
        memcpy(o->ro_old, o->ro, n); memcpy(o->ru_old, o->ru, n); memcpy(o->rv_old, o->rv, n); memcpy(o->re_old, o->re, n);
        stage(o);
        stage(o);
        for (int c = 0; c < o->nc; c++) {
            if ((o->flag[c] & CFD2D_CELL_FLAG_LIM) > 0) continue;
            o->ro[c] = 0.5 * (o->ro_old[c] + o->ro[c]);
            o->ru[c] = 0.5 * (o->ru_old[c] + o->ru[c]);
            o->rv[c] = 0.5 * (o->rv_old[c] + o->rv[c]);
            o->re[c] = 0.5 * (o->re_old[c] + o->re[c]);
            Param par;
            prim(o, c, &par);
            if (par.r < o->lim[0]) o->flag[c] |= CFD2D_CELL_FLAG_LIM;
            if (par.r > o->lim[1]) o->flag[c] |= CFD2D_CELL_FLAG_LIM;
            if (par.p < o->lim[2]) o->flag[c] |= CFD2D_CELL_FLAG_LIM;
            if (par.p > o->lim[3]) o->flag[c] |= CFD2D_CELL_FLAG_LIM;
            if (fabs(par.u) > o->lim[4]) o->flag[c] |= CFD2D_CELL_FLAG_LIM;
            if (fabs(par.v) > o->lim[4]) o->flag[c] |= CFD2D_CELL_FLAG_LIM;
        }
        remediate(o);
        if (o->err) return o->err;
    }
    return 0;
}

void fvm_oracle_calc_grad(fvm_oracle* o, double* grad8) {
    calc_grad(o);
    memcpy(grad8, o->grad, (size_t)o->nc * 64);
}

void fvm_oracle_edge_fluxes(fvm_oracle* o, double* flux4) {
    for (int e = 0; e < o->ne; e++) edge_flux(o, e, flux4 + 4 * (size_t)e);
}

void fvm_oracle_get_primitive(fvm_oracle* o, double* r, double* p, double* T, double* u, double* v, double* cz) {
    for (int c = 0; c < o->nc; c++) {
        Param q;
        prim(o, c, &q);
        if (r) r[c] = q.r; if (p) p[c] = q.p; if (T) T[c] = q.T; if (u) u[c] = q.u; if (v) v[c] = q.v; if (cz) cz[c] = q.cz;
    }
}

long long fvm_oracle_newton_iters(const fvm_oracle* o) { return o->newton_iters; }
long long fvm_oracle_riemann_calls(const fvm_oracle* o) { return o->riemann_calls; }

int fvm_oracle_rim_orig(int n, const double* in8, double gam, int max_newton, double* out5, int32_t* iters) {
    int bad = 0;
    if (max_newton <= 0) max_newton = 1000;
    for (int i = 0; i < n; i++) {
        const double* a = in8 + 8 * (size_t)i;
        double* q = out5 + 5 * (size_t)i;
        int it = riemann(&q[0], &q[1], &q[2], &q[3], &q[4], a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], gam, max_newton);
        if (iters) iters[i] = it;
        if (it < 0) bad = CFD2D_ENEWTON;
    }
    return bad;
}

void fvm_oracle_calc_flux(int n, const double* in12, double gam, int flux, double* out4) {
    for (int i = 0; i < n; i++) {
        const double* a = in12 + 12 * (size_t)i;
        double* q = out4 + 4 * (size_t)i;
        Param L, R;
        memset(&L, 0, sizeof L); memset(&R, 0, sizeof R);
        L.r = a[0]; L.p = a[1]; L.u = a[2]; L.v = a[3]; L.E = a[4];
        R.r = a[5]; R.p = a[6]; R.u = a[7]; R.v = a[8]; R.E = a[9];
        if (flux == CFD2D_FLUX_LAX) flux_lax(&q[0], &q[1], &q[2], &q[3], &L, &R, a[10], a[11], gam);
        else flux_godunov(&q[0], &q[1], &q[2], &q[3], &L, &R, a[10], a[11], gam, 1000);
    }
}
