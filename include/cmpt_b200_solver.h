/* cmpt_b200_solver.h — C binding of the solver classes of include/cmpt/eigen_ex/{lanczos,arnoldi}.hpp.
 *
 * The reference's user-facing API is a pair of C++ class templates (LanczosEigenSolver lanczos.hpp:468-927,
 * ArnoldiEigenSolver arnoldi.hpp:444-1027).  C++ users include the drop-in headers directly; every other host
 * language (the Python tests and bench.py use ctypes) drives the same classes through this binding, which
 * adds nothing of its own: each function forwards to the member of the same name.
 *
 * kind: CMBS_LANCZOS, CMBS_ARNOLDI or CMBS_THICK_RESTART (integer settings wanted / maxBasis / keep / maxRestarts); dtype: CMB_F64 or CMB_C64 (Scalar = double / std::complex<double>).
 * Vectors are contiguous arrays of dtype elements; matrices are column-major.  Lanczos eigenvalues are double,
 * Arnoldi eigenvalues/eigenvectors are complex (interleaved re,im) for both dtypes.
 */
#ifndef CMPT_B200_SOLVER_H_
#define CMPT_B200_SOLVER_H_

#include <stddef.h>

#include "cmpt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

typedef struct cmbs_solver cmbs_solver;
enum {
  CMBS_LANCZOS = 0,
  CMBS_ARNOLDI = 1,
  CMBS_THICK_RESTART = 2,         /* additive: ThickRestartLanczos (thick_restart.hpp) */
  CMBS_THICK_RESTART_ARNOLDI = 3  /* additive: ThickRestartArnoldi (arnoldi_restart.hpp); integer setting "which": 0 largest |lambda|, 1 largest real, 2 smallest real */
};

int cmbs_create(int kind, cmb_dtype dtype, cmbs_solver** out);
int cmbs_destroy(cmbs_solver* s);

/* setMatrixMultiplication(DeviceOperator): op stays owned by the caller and must outlive the solver's use of it */
int cmbs_set_operator(cmbs_solver* s, cmb_ctx* ctx, cmb_op* op);
/* setMatrixMultiplication(std::function, height): the legacy host-callback path */
int cmbs_set_callback(cmbs_solver* s, int64_t height, cmb_matmul_fn fn, void* user);
/* integer settings: "minIterations" "maxIterations" "maxEigenvalues" "reorthogonalizeInterval" (Lanczos)
 * "computeEigenvectorsOn" "reserveSize"; real settings: "tolerance" "threshold" "eigenvalueShift"
 * (Arnoldi's shift is complex: use cmbs_set_complex) */
int cmbs_set_int(cmbs_solver* s, const char* name, int64_t value);
int cmbs_set_real(cmbs_solver* s, const char* name, double value);
int cmbs_set_complex(cmbs_solver* s, const char* name, double re, double im);
int cmbs_get_int(cmbs_solver* s, const char* name, int64_t* value);
int cmbs_get_real(cmbs_solver* s, const char* name, double* value);
int cmbs_set_indices_for_convergence(cmbs_solver* s, const int64_t* idx, int64_t n);
/* setInitialVector(v); n == 0 selects the default random vector (std::mt19937 default seed) */
int cmbs_set_initial_vector(cmbs_solver* s, const void* v, int64_t n);
int cmbs_set_orthogonalizing_vectors(cmbs_solver* s, int64_t nvec, const void* vecs, int64_t ld);

int cmbs_compute(cmbs_solver* s);
int cmbs_continue_to_compute(cmbs_solver* s);
int cmbs_compute_with_restarts(cmbs_solver* s, int64_t cycles); /* Arnoldi only (additive) */
int cmbs_clear(cmbs_solver* s);
int cmbs_clear_computed_data(cmbs_solver* s);

/* results: sizes through cmbs_get_int ("iterations" "nvectors" "neigenvalues" "nalpha" "nbeta" "nlog"
 * "hasWARN" "hasERROR" "hessenbergSize" "matrixHeight") */
int cmbs_get_eigenvalues(cmbs_solver* s, void* out);           /* double[nev] (Lanczos) / complex[nev] (Arnoldi) */
int cmbs_eigenvectors_ptr(cmbs_solver* s, const void** ptr, int64_t* rows, int64_t* cols); /* no copy */
int cmbs_get_ritz_residuals(cmbs_solver* s, double* out);
int cmbs_get_alpha_beta(cmbs_solver* s, double* alpha, double* beta);     /* Lanczos */
int cmbs_get_hessenberg(cmbs_solver* s, void* out);                        /* Arnoldi: dtype, size^2 col-major */
int cmbs_get_residue(cmbs_solver* s, double* out);                         /* Arnoldi */
int cmbs_get_small_eigenvectors(cmbs_solver* s, void* out, int64_t* rows, int64_t* cols); /* es_tri / eigenvectors_h */
int cmbs_get_basis_vector(cmbs_solver* s, int64_t k, void* out);          /* lanczosvectors()[k] / arnoldivectors()[k] */
int cmbs_get_log_line(cmbs_solver* s, int64_t i, char* buf, int64_t buflen);
/* convergenceLog()[index]: *n receives the series length; out (may be NULL) receives the values
 * (double per entry for Lanczos, complex for Arnoldi) */
int cmbs_get_convergence_log(cmbs_solver* s, int64_t index, void* out, int64_t* n);
double cmbs_device_bytes(cmbs_solver* s);

/* LanczosExponentialSolver (lanczos.hpp:1002-1164), Lanczos solvers only: out = exp(x A) v.
 * solve_with_lanczos: es.compute() then the eigen-expansion with the solver's initial vector (local slab out);
 * solve_with_taylor: Taylor expansion through the solver's matrixMultiplication() (auto_division != 0 splits x). */
int cmbs_exp_solve_with_lanczos(cmbs_solver* s, double x_re, double x_im, void* out);
int cmbs_exp_solve_with_taylor(cmbs_solver* s, double x_re, double x_im, double matrix_radius, int auto_division,
                               const void* in, void* out);

/* host-side Ritz solvers, exposed for testing against LAPACK (no GPU needed) */
int cmbs_host_tridiagonal_eigen(int64_t n, const double* alpha, const double* beta, double* w, double* z /*nullable*/);
int cmbs_host_hessenberg_eigen(int64_t n, const void* h_complex, void* w_complex, void* v_complex /*nullable*/);
/* general complex matrix (column-major): Householder reduction to Hessenberg form, then the solver above; the projected
 * matrix of a thick-restarted Arnoldi run is not Hessenberg any more */
int cmbs_host_general_eigen(int64_t n, const void* a_complex, void* w_complex, void* v_complex /*nullable*/);
/* dense real symmetric n x n (column-major), cyclic Jacobi (detail/symmetric_eigen.hpp): the projected matrix of the
 * thick-restart driver.  w ascending, z column-major eigenvectors. */
int cmbs_host_symmetric_eigen(int64_t n, const double* a, double* w, double* z);

/* pinned host memory for staging inputs/results at full PCIe rate */
int cmb_host_alloc(size_t bytes, void** out);
int cmb_host_free(void* p);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
