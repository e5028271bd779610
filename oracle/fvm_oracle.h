/* fvm_oracle.h -- TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's FVM_TVD path.
 * See fvm_oracle.c for the file:line map.  Uses the product's plain-C mesh/phys/ctrl structs so
 * the checker and the thing checked consume byte-identical inputs.  Serial (global mesh) only. */
#ifndef FVM_ORACLE_H
#define FVM_ORACLE_H
#include "../include/cfd2d_fvm.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct fvm_oracle fvm_oracle;
fvm_oracle* fvm_oracle_create(const cfd2d_mesh* m, const cfd2d_phys* p, const cfd2d_ctrl* c);
void   fvm_oracle_destroy(fvm_oracle* o);
void   fvm_oracle_set_state(fvm_oracle* o, const double* ro, const double* ru, const double* rv,
                            const double* re, const uint32_t* flag);
double fvm_oracle_calc_time_step(fvm_oracle* o);
int    fvm_oracle_step(fvm_oracle* o, int nsteps);           /* 0 or CFD2D_ENEWTON */
void   fvm_oracle_get_state(fvm_oracle* o, double* ro, double* ru, double* rv, double* re,
                            double* cTau, uint32_t* flag);
void   fvm_oracle_calc_grad(fvm_oracle* o, double* grad8);
void   fvm_oracle_edge_fluxes(fvm_oracle* o, double* flux4);
void   fvm_oracle_get_primitive(fvm_oracle* o, double* r, double* p, double* T, double* u, double* v, double* cz);
long long fvm_oracle_newton_iters(const fvm_oracle* o);       /* total Newton iterations so far */
long long fvm_oracle_riemann_calls(const fvm_oracle* o);
int    fvm_oracle_rim_orig(int n, const double* in8, double gam, int max_newton, double* out5, int32_t* iters);
void   fvm_oracle_urs(int n, double M, double Cp, int mode, double* io8);   /* Material::URS, in place */
void   fvm_oracle_calc_flux(int n, const double* in12, double gam, int flux, double* out4);
#ifdef __cplusplus
}
#endif
#endif
