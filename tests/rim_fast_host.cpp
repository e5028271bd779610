// TEST INFRASTRUCTURE ONLY: host build of cfd-2d_b200/csrc/fvm_riemann_fast.cuh (CFD2D_RIM_HOST) so
// that its algebra -- shared reciprocals, the x^(1/7) Newton kernel, the one-division shock star --
// can be property-tested on CPU against the oracle before any GPU time is spent.  The product never
// links this file.
#define CFD2D_RIM_HOST 1
#include <cmath>
using std::exp; using std::log; using std::sqrt; using std::fabs;
#include "../cfd-2d_b200/csrc/fvm_riemann_fast.cuh"

static RimC make_rim_host(double GAM) {   // the expressions of rim_orig's constants, global.cpp:235-249
    RimC k;
    k.GAM = GAM; k.AGAM = GAM - 1.0; k.DGAM = 2.0 / k.AGAM; k.GGAM = sqrt(GAM * k.AGAM); k.HGAM = k.AGAM / 2.0;
    k.FGAM = 3.0 * GAM - 1.0; k.OGAM = k.AGAM / (2.0 * GAM); k.QGAM = GAM + 1.0; k.PGAM = k.QGAM / (2.0 * GAM);
    k.RGAM = 4.0 * GAM; k.SGAM = GAM * k.AGAM; k.TGAM = k.QGAM / 2.0; k.IAGAM = 1 / k.AGAM; k.DG1 = 1 + k.DGAM;
    k.DGGG = k.DGAM * k.GGAM; k.ISGAM = 1.0 / k.SGAM; k.IDG1 = 1.0 / k.DG1;
    return k;
}

extern "C" void rim_fast_host(int n, const double* in8, int max_newton, double* out5, int* iters) {
    RimC k = make_rim_host(1.4);
    for (int i = 0; i < n; i++) {
        const double* a = in8 + 8 * (long)i;
        double* q = out5 + 5 * (long)i;
        iters[i] = rim_orig_fast(k, max_newton, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], q[0], q[1], q[2], q[3], q[4]);
    }
}

extern "C" void flux_fast_host(int n, const double* in12, double* out4) {
    RimC k = make_rim_host(1.4);
    for (int i = 0; i < n; i++) {
        const double* a = in12 + 12 * (long)i;
        Prim L = {a[0], a[1], a[2], a[3]}, R = {a[5], a[6], a[7], a[8]};
        double* q = out4 + 4 * (long)i;
        flux_godunov_fast(k, 1000, L, R, a[10], a[11], q[0], q[1], q[2], q[3]);
    }
}

extern "C" void pow17_host(int n, const double* x, double* y) {
    for (int i = 0; i < n; i++) y[i] = pow17(x[i], 0.4 / 2.8);
}
