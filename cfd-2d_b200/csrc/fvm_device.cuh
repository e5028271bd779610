// fvm_device.cuh -- device-side physics of the FVM_TVD path (FP64, sm_100a).
//
// Every function states the reference code whose arithmetic it reproduces.  The whole library is
// compiled with -fmad=false: the reference x86-64 build has no FMA contraction, and +,-,*,/,sqrt
// are IEEE-exact on both sides, so apart from exp/log (<= 1 ulp on both, not identical) results
// agree bit for bit.  Expressions keep the reference's left-to-right association.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CFD2D_GR 8.314472   // Material::gR, reference src/global.cpp:6

// Material constants derived once on the host with the reference's own expressions
// (Material::URS, global.cpp:11-12: Cv = Cp - gR/M; gam = Cp/Cv).
struct MatC { double M, Cv, gam, gm1; };

// Constants of rim_orig (global.cpp:235-249) for the fixed GAM = 1.4 the flux loop passes
// (fvm_tvd.cpp:345,398).  Pure functions of GAM, evaluated on the host by the same expressions.
struct RimC {
    double GAM, AGAM, DGAM, GGAM, HGAM, FGAM, OGAM, QGAM, PGAM, RGAM, SGAM, TGAM;
    double IAGAM;   // (1/AGAM)            global.cpp:389,395
    double DG1;     // (1+DGAM)            global.cpp:384,391
    double DGGG;    // DGAM*GGAM           global.cpp:384,391
    double ISGAM;   // 1/SGAM, 1/DG1: reciprocals of the constant divisors, used only by the reduced-instruction
    double IDG1;    //                solver (fvm_riemann_fast.cuh); the bit-faithful rim_orig_dev divides as written
};

struct Prim { double r, p, u, v; };

// Material::URS (global.cpp:9-30), the three expressions the path uses
__device__ __forceinline__ double urs_p(double r, double e, double gm1) { return r * e * gm1; }            // mode 0, :16
__device__ __forceinline__ double urs_e(double p, double r, double gm1) { return p / (r * gm1); }          // mode 1, :21
__device__ __forceinline__ double urs_r(double p, double T, double M) { return p * M / (T * CFD2D_GR); }   // mode 2, :26

// FVM_TVD::convertConsToPar (fvm_tvd.cpp:803-813) + Material::URS mode 0 (global.cpp:15-18).
// Only r,p,u,v are kept: reconstruct/calcFlux read nothing else of an inner state.
__device__ __forceinline__ Prim cons_to_prim(double ro, double ru, double rv, double re, double gm1) {
    Prim w;
    w.r = ro;
    w.u = ru / ro;
    w.v = rv / ro;
    double E = re / ro;
    double e = E - 0.5 * (w.u * w.u + w.v * w.v);
    w.p = urs_p(w.r, e, gm1);
    return w;
}

// T as convertConsToPar leaves it: URS(1) after URS(0): e = p/(r*(gam-1)); T = e/Cv (global.cpp:20-23)
__device__ __forceinline__ double prim_T(const Prim& w, const MatC& m) {
    double e = urs_e(w.p, w.r, m.gm1);
    return e / m.Cv;
}

// cz of URS(0): sqrt(gam*p/r) (global.cpp:17)
__device__ __forceinline__ double prim_cz(const Prim& w, const MatC& m) { return sqrt(m.gam * w.p / w.r); }

// FVM_TVD::boundaryCond (fvm_tvd.cpp:694-711) with CFDBnd{Inlet,Outlet,WallSlip}::run
// (bnd_cond.cpp:75-110).  pL = (possibly extrapolated) r,p,u,v; TL = the CELL-CENTRE temperature
// (reconstruct never updates T: SURVEY Q3).  Returns the ghost r,p,u,v and, for the LF flux, E.
__device__ __forceinline__ Prim ghost_state(const Prim& pL, double TL, int kind, const double* __restrict__ par,
                                            double nx, double ny, const MatC& m, double* E_out) {
    Prim pR;
    double T;
    if (kind == 1) {            // CFDBndInlet
        pR.u = par[0]; pR.v = par[1]; T = par[2]; pR.p = par[3];
    } else if (kind == 2) {     // CFDBndOutlet
        pR.u = pL.u; pR.v = pL.v; T = TL; pR.p = pL.p;
    } else {                    // CFDBndWallSlip (and "no-slip")
        double Un = pL.u * nx + pL.v * ny;
        double Vx = nx * Un * 2.0;
        double Vy = ny * Un * 2.0;
        pR.u = pL.u - Vx; pR.v = pL.v - Vy; T = TL; pR.p = pL.p;
    }
    pR.r = urs_r(pR.p, T, m.M);                       // URS(2), global.cpp:26
    if (E_out) {
        double e = urs_e(pR.p, pR.r, m.gm1);          // URS(1), global.cpp:21
        *E_out = e + 0.5 * (pR.u * pR.u + pR.v * pR.v); // fvm_tvd.cpp:703
    }
    return pR;
}

// FVM_TVD::reconstruct (fvm_tvd.cpp:646-691): q += gq.x*DL.x + gq.y*DL.y.  FUSED = false keeps the
// reference's rounding (mul, mul, add, add -- the Lax-Friedrichs variants and the order-faithful
// Riemann statement are bit-exact through it); FUSED = true (only the reduced-instruction Godunov
// path, which is held to 1e-12 and not to the bit) forms the same sum with two FMAs.
template <bool FUSED>
__device__ __forceinline__ double recon1(double q, double gx, double gy, double dx, double dy) {
    if (FUSED) return fma(gy, dy, fma(gx, dx, q));   // explicit fma(): fused on the device whatever -fmad says
    return q + (gx * dx + gy * dy);
}

// One side of the Newton function of rim_orig (global.cpp:281-304): given the trial pressure P
// returns F and its derivative FS for the side with state (PS, CS, RCS).
__device__ __forceinline__ void rim_side(const RimC& k, double P, double PS, double CS, double RCS,
                                         double& F, double& FS) {
    double PP = P / PS;
    if (PS > P) {                                  // rarefaction: lbl1 / lbl3
        double ZF = CS * exp(log(PP) * k.OGAM);
        F = k.DGAM * (ZF - CS);
        FS = ZF / (k.GAM * P);
    } else {                                       // shock
        double PK = k.PGAM * PP + k.OGAM;
        double ZN = RCS * sqrt(PK);
        F = (P - PS) / ZN;
        FS = (k.QGAM * PP + k.FGAM) / (k.RGAM * ZN * PK);
    }
}

// One side of the Riemann problem in the edge frame: s = -1 for the left state "B", +1 for the right
// state "E".  The reference writes the two sides out separately with opposite signs
// (global.cpp:318-349); x + (-1)*y == x - y exactly, so one signed routine reproduces both.
struct RimSide { double R, P, U, V, C, RC, s; };
struct RimWave { bool rar; double ZD; double Ustar; double head, tail; };

// wave speeds of one side for the converged P (global.cpp:318-349)
__device__ __forceinline__ void rim_waves(const RimC& k, double P, const RimSide& S, RimWave& w) {
    w.rar = S.P > P;
    if (w.rar) {                                   // lbl6 / lbl8: rarefaction
        double ZF = S.C * exp(log(P / S.P) * k.OGAM);
        w.Ustar = S.U - S.s * (k.DGAM * (S.C - ZF));
        w.head = S.U + S.s * S.C;
        w.tail = w.Ustar + S.s * ZF;
        w.ZD = ZF;
    } else {                                       // shock moving with speed D
        double D = S.U + S.s * sqrt((k.TGAM * P + k.HGAM * S.P) / S.R);
        w.Ustar = 0.0;
        w.head = D;
        w.tail = D;
        w.ZD = D;
    }
}

// density and velocity behind a shock (global.cpp:321-324 / :338-341)
__device__ __forceinline__ void rim_shock_star(double P, const RimSide& S, double D, double& Rst, double& Ust) {
    double UD = S.U - D;
    double RUD = S.R * UD;
    Rst = RUD * RUD / (S.P - P + RUD * UD);
    Ust = D + RUD / Rst;
}

// rim_orig (global.cpp:232-405) with WB = WE = 0.  Quantities the reference computes but never
// reads on the taken path (e.g. ZFB on a shock side, global.cpp:315; the star state of the side the
// sampling does not select) are skipped; every value that is used is formed by the reference's own
// expression.  Branch coherence: both the Newton function and the wave speeds are evaluated for the
// HIGHER-pressure side first -- whenever the star pressure lies between PB and PE (all smooth-flow
// edges) every lane of a warp then takes the rarefaction branch in the first call and the shock
// branch in the second, whatever the edge orientation; results are mapped back to (B, E) before any
// order-sensitive arithmetic.  Returns the Newton iteration count, or -1 if `max_newton` was
// reached (the reference has no cap and would hang).
__device__ __forceinline__ int rim_orig_dev(const RimC& k, int max_newton,
                                            double RB, double PB, double UB, double VB,
                                            double RE, double PE, double UE, double VE,
                                            double& RI, double& EI, double& PI, double& UI, double& VI) {
    const double eps = 1.0e-5;
    RimSide B, E;
    B.R = RB; B.P = PB; B.U = UB; B.V = VB; B.s = -1.0;
    E.R = RE; E.P = PE; E.U = UE; E.V = VE; E.s = 1.0;
    B.C = sqrt(k.GAM * PB / RB);
    E.C = sqrt(k.GAM * PE / RE);
    B.RC = RB * B.C;
    E.RC = RE * E.C;
    double DU = UB - UE;
    double P = 0.0;
    RimWave wB, wE;
    bool vacuum = false;
    int it = 0;
    if (DU < -2.0 * (B.C + E.C) / k.AGAM) {        // vacuum, global.cpp:265-276 (RF=RS=EF=ES=UF=US=0)
        vacuum = true;
        wB.rar = wE.rar = false; wB.ZD = wE.ZD = 0.0; wB.Ustar = wE.Ustar = 0.0;
        wB.head = UB - B.C;                        // SBL
        wB.tail = UB + 2.0 * B.C / k.AGAM;         // SFL
        wE.tail = UE - 2.0 * E.C / k.AGAM;         // SSL
        wE.head = UE + E.C;                        // SEL
    } else {
        const bool sw = PE > PB;                   // hi = E when the right pressure is larger
        const RimSide hi = sw ? E : B, lo = sw ? B : E;
        P = (PB * E.RC + PE * B.RC + DU * B.RC * E.RC) / (B.RC + E.RC);   // global.cpp:277
        for (;;) {
            if (P < eps) P = eps;
            double Fh, FSh, Fl, FSl;
            rim_side(k, P, hi.P, hi.C, hi.RC, Fh, FSh);
            rim_side(k, P, lo.P, lo.C, lo.RC, Fl, FSl);
            double F1 = sw ? Fl : Fh, F2 = sw ? Fh : Fl;          // back to (B, E) order
            double FS1 = sw ? FSl : FSh, FS2 = sw ? FSh : FSl;
            double res = DU - F1 - F2;
            double DP = res / (FS1 + FS2);
            P = P + DP;
            ++it;
            if (!(fabs(res) > eps)) break;
            if (it >= max_newton) { it = -1; break; }
        }
        RimWave wh, wl;
        rim_waves(k, P, hi, wh);
        rim_waves(k, P, lo, wl);
        wB = sw ? wl : wh;
        wE = sw ? wh : wl;
    }
    const double SBL = wB.head, SFL = wB.tail, SSL = wE.tail, SEL = wE.head;
    // sampling at x/t = 0, global.cpp:353-400
    if (SEL <= 0.0) {
        RI = RE; EI = E.C * E.C / k.SGAM; UI = UE; VI = VE;          // EE, global.cpp:260
    } else if (SBL >= 0.0) {
        RI = RB; EI = B.C * B.C / k.SGAM; UI = UB; VI = VB;          // EB, global.cpp:259
    } else if ((SSL >= 0.0) && (SFL <= 0.0)) {
        double RS = 0.0, US = wE.Ustar;
        if (!vacuum && !wE.rar) rim_shock_star(P, E, wE.ZD, RS, US);   // US decides the side
        if (US >= 0.0) {                           // left star state
            double RF = 0.0, EF = 0.0, UF = wB.Ustar;
            if (!vacuum) {
                if (wB.rar) {
                    EF = wB.ZD * wB.ZD / k.SGAM;
                    RF = P / (k.AGAM * EF);
                } else {
                    rim_shock_star(P, B, wB.ZD, RF, UF);
                    EF = P / (k.AGAM * RF);
                }
            }
            RI = RF; EI = EF; UI = UF; VI = VB;
        } else {                                   // right star state
            double ES;
            if (wE.rar) {
                ES = wE.ZD * wE.ZD / k.SGAM;
                RS = P / (k.AGAM * ES);
            } else {
                ES = P / (k.AGAM * RS);
            }
            RI = RS; EI = ES; UI = US; VI = VE;
        }
    } else if (SFL > 0.0) {
        double EB = B.C * B.C / k.SGAM;
        UI = (UB + k.DGGG * sqrt(EB)) / k.DG1;
        VI = VB;
        EI = (UI * UI) / k.SGAM;
        RI = RB * exp(log(EI / EB) * k.IAGAM);
    } else {
        double EE = E.C * E.C / k.SGAM;
        UI = (UE - k.DGGG * sqrt(EE)) / k.DG1;
        VI = VE;
        EI = (UI * UI) / k.SGAM;
        RI = RE * exp(log(EI / EE) * k.IAGAM);
    }
    PI = k.AGAM * EI * RI;
    return it;
}

// FVM_TVD::calcFlux, Godunov block (fvm_tvd.cpp:604-622).
__device__ __forceinline__ int flux_godunov_dev(const RimC& k, int max_newton, const Prim& L, const Prim& R,
                                                double nx, double ny, double& fr, double& fu, double& fv, double& fe) {
    double unl = L.u * nx + L.v * ny;
    double unr = R.u * nx + R.v * ny;
    double utl = L.u * ny - L.v * nx;
    double utr = R.u * ny - R.v * nx;
    double RI, EI, PI, UN, UT;
    int it = rim_orig_dev(k, max_newton, L.r, L.p, unl, utl, R.r, R.p, unr, utr, RI, EI, PI, UN, UT);
    double UI = UN * nx + UT * ny;
    double VI = UN * ny - UT * nx;
    fr = RI * UN;
    fu = fr * UI + PI * nx;
    fv = fr * VI + PI * ny;
    fe = (RI * (EI + 0.5 * (UI * UI + VI * VI)) + PI) * UN;
    return it;
}

// The commented Lax-Friedrichs block of FVM_TVD::calcFlux (fvm_tvd.cpp:623-642); EL/ER are the
// total specific energies pL.E / pR.E (cell-centre value on an inner side, ghost value on a boundary).
__device__ __forceinline__ void flux_lax_dev(double GAM, const Prim& L, double EL, const Prim& R, double ER,
                                             double nx, double ny, double& fr, double& fu, double& fv, double& fe) {
    double unl = L.u * nx + L.v * ny;
    double unr = R.u * nx + R.v * ny;
    double al = fabs(unl) + sqrt(GAM * L.p / L.r);
    double ar = fabs(unr) + sqrt(GAM * R.p / R.r);
    double alpha = (al > ar) ? al : ar;            // _max_, grid.h:102
    double rol = L.r, rul = L.r * L.u, rvl = L.r * L.v, rel = L.r * EL;
    double ror = R.r, rur = R.r * R.u, rvr = R.r * R.v, rer = R.r * ER;
    double frl = rol * unl;
    double frr = ror * unr;
    fr = 0.5 * (frr + frl - alpha * (ror - rol));
    fu = 0.5 * (frr * R.u + frl * L.u + (R.p + L.p) * nx - alpha * (rur - rul));
    fv = 0.5 * (frr * R.v + frl * L.v + (R.p + L.p) * ny - alpha * (rvr - rvl));
    fe = 0.5 * ((rer + R.p) * unr + (rel + L.p) * unl - alpha * (rer - rel));
}
