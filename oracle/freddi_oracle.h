/* TEST INFRASTRUCTURE — C interface of the CPU restatement (oracle/freddi_oracle.cpp).
 * Same entry points as oracle/ref_driver.cpp (prefix orc_ instead of ref_) so that tests can swap them.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this library. */
#ifndef FREDDI_ORACLE_H
#define FREDDI_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/freddi_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
void* orc_create(const fb200_args_t* a, const double* lambdas, int n_lambdas, int* err_code, char* err, size_t err_len);
void orc_destroy(void* h);
int orc_n_cols(void* h, int cold_disk);
int orc_step(void* h); /* 0 stepped, 1 RadiusCollapse */
long orc_last_picard(void* h);
void orc_row(void* h, int cold_disk, double* out);
void orc_state(void* h, double* t, int64_t* i_t, int64_t* first, int64_t* last);
void orc_info(void* h, fb200_model_info_t* out);
int orc_field(void* h, int field, double* out);
double orc_flux(void* h, int region, double lambda);
double orc_mdot_wind(void* h);
int64_t orc_run(const fb200_args_t* a, const double* lambdas, int n_lambdas, int cold_disk, double* rows, int64_t rows_cap,
                double* Fsnap, int64_t* first, int64_t* last, int64_t* picard, char* err, size_t err_len);
int64_t orc_run_ensemble(const fb200_args_t* args, int64_t n_models, const double* lambdas, int n_lambdas, int cold_disk,
                         int n_threads, int64_t max_steps, double* rows, int64_t rows_per_model, int32_t* rows_valid,
                         double* construct_seconds, double* evolve_seconds);
/* Building blocks exposed for unit tests */
double orc_planck_nu1_nu2(double T, double nu1, double nu2, double tol, long* trial_steps);
double orc_trapz(const double* x, const double* y, size_t first, size_t last);
double orc_T_GR(double r1, double ak, double Mx, double Mdot);
void orc_solver_step(double tau, double eps, double left_bc, double right_bc, const double* A, const double* B, const double* C,
                     double one_minus_m, double n, double D, const double* x, double* y, size_t first, size_t last, long* iters);
#ifdef __cplusplus
}
#endif
#endif
