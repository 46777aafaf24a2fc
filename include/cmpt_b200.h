/* cmpt_b200.h — C-ABI of libcmpt_b200.so: the device side of cmpt-eigenex's Lanczos/Arnoldi path.
 *
 * The reference (versmc/cmpt-eigenex) is header-only C++ with no FFI; its only plug-in boundary is
 *     using MatMulFunction = std::function<void(const Scalar*, Scalar*)>;   lanczos.hpp:116, arnoldi.hpp:65
 * plus the solver classes built on it.  This header is the boundary the B200 build puts under those
 * classes: plain pointers and sizes, opaque handles, int status codes, no C++/torch types.  The C++
 * drop-in headers (include/cmpt/eigen_ex/lanczos.hpp, arnoldi.hpp) and the ctypes binding
 * (cmpt_eigenex_b200/capi.py) are both written against exactly these entry points.
 *
 * Conventions
 *   - every function returns CMB_OK (0) or a negative error class and never throws;
 *     cmb_last_error() gives the message of the calling thread's last failure;
 *   - dtype CMB_F64 = double, CMB_C64 = std::complex<double> (interleaved re,im);
 *   - a context is one GPU (one rank).  In a distributed context every vector/basis/operator is
 *     row-partitioned: this rank owns global rows [row_begin,row_end); host vectors passed in or
 *     out are always the LOCAL slab (row_end-row_begin entries);
 *   - host arrays are copied during the call; nothing retains caller pointers except the legacy
 *     callback operator (fn + user).
 */
#ifndef CMPT_B200_H_
#define CMPT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

typedef struct cmb_ctx cmb_ctx;       /* one GPU / one rank: streams, workspaces, NCCL communicator      */
typedef struct cmb_op cmb_op;         /* operator shard resident in HBM (replaces the matmul callback)    */
typedef struct cmb_krylov cmb_krylov; /* device-resident Krylov state (LanczosBase / ArnoldiBase members) */

typedef enum { CMB_F64 = 0, CMB_C64 = 1 } cmb_dtype;

enum {
  CMB_OK = 0,
  CMB_ERR_INVALID = -1,     /* bad argument / wrong state                       */
  CMB_ERR_CUDA = -2,        /* CUDA runtime or driver error                      */
  CMB_ERR_NOMEM = -3,       /* host or device allocation failed                  */
  CMB_ERR_NCCL = -4,        /* NCCL missing or failing                           */
  CMB_ERR_UNSUPPORTED = -5, /* valid request this build does not implement       */
  CMB_ERR_NO_DEVICE = -6    /* no usable CUDA device (the product has no CPU fallback) */
};

/* step status bits returned by the step functions */
enum {
  CMB_STEP_OK = 0,        /* a new basis vector was produced                                          */
  CMB_STEP_BREAKDOWN = 1, /* beta / residue <= threshold: vector dropped (lanczos.hpp:433-437)        */
  CMB_STEP_NOSTART = 2,   /* start vector norm < threshold: basis stays empty (lanczos.hpp:316-318)   */
  CMB_STEP_FULL = 4       /* basis already spans the whole space (lanczos.hpp:332)                    */
};

/* ---- context --------------------------------------------------------------------------------- */
const char* cmb_version(void);
const char* cmb_last_error(void);
int cmb_device_count(int* count);
int cmb_ctx_create(int device, cmb_ctx** out);
/* nccl_id: the 128 bytes of an ncclUniqueId made by cmb_nccl_unique_id() on rank 0 and broadcast by
 * the host program (torch.distributed, MPI, ...). */
int cmb_nccl_unique_id(void* id128);
int cmb_ctx_create_dist(int device, int rank, int nranks, const void* nccl_id, cmb_ctx** out);
int cmb_ctx_destroy(cmb_ctx* ctx);
int cmb_ctx_rank(const cmb_ctx* ctx);
int cmb_ctx_nranks(const cmb_ctx* ctx);
int cmb_ctx_sync(cmb_ctx* ctx);
/* CUDA-event timer on the context's compute stream (the stream every kernel of this library is
 * launched on) and the number of kernels launched through this context so far. */
int cmb_ctx_timer_start(cmb_ctx* ctx);
int cmb_ctx_timer_stop(cmb_ctx* ctx, double* milliseconds);
uint64_t cmb_ctx_launch_count(const cmb_ctx* ctx);
/* per-kernel-family accumulated device time; enabling inserts events around every launch */
int cmb_ctx_profile(cmb_ctx* ctx, int enable);
int cmb_ctx_profile_get(cmb_ctx* ctx, const char* family, double* total_ms, uint64_t* launches);
/* write a buffer larger than L2 (flushes it between timed iterations) */
int cmb_ctx_flush_l2(cmb_ctx* ctx);

/* ---- operators: replace `matrixMultiplication_` (lanczos.hpp:154,178-189; arnoldi.hpp:107,131-140) ---
 * CSR shard: rows [row_begin,row_end) of an n_global x n_global matrix; rowptr has
 * (row_end-row_begin+1) entries starting at 0; col holds GLOBAL column indices.  The shard is converted
 * on the device to SELL-32 and, in a distributed context, its halo lists are built. */
int cmb_op_csr_create(cmb_ctx* ctx, cmb_dtype dtype, int64_t n_global, int64_t row_begin, int64_t row_end,
                      const int64_t* rowptr, const int32_t* col, const void* val, cmb_op** out);
/* the uniform row partition every distributed object uses: rank q owns [begin(q), begin(q+1)) */
int64_t cmb_partition_begin(int64_t n_global, int nranks, int rank);
/* host-only (no GPU): halo plan of a CSR shard under that partition.  col_local (nnz entries, may be NULL)
 * receives the remapped column indices: owned columns -> [0,n_local), remote ones -> n_local + position in the
 * sorted list of distinct remote columns; halo_cols (capacity entries, may be NULL) receives that list. */
int cmb_plan_halo(int64_t n_global, int nranks, int rank, int64_t nnz, const int32_t* col, int32_t* col_local,
                  int64_t* halo_count, int64_t* per_owner_counts, int32_t* halo_cols, int64_t halo_capacity);
/* dense row-major rows [row_begin,row_end) x n_global (cfg 1) */
int cmb_op_dense_create(cmb_ctx* ctx, cmb_dtype dtype, int64_t n_global, int64_t row_begin, int64_t row_end,
                        const void* a_rows, cmb_op** out);
/* matrix-free spin-1/2 Heisenberg chain (cfg 5): H = J sum_i [SzSz + (S+S- + S-S+)/2]_{i,i+1}; the
 * top log2(nranks) bits of the state index are the rank. */
int cmb_op_heisenberg_create(cmb_ctx* ctx, cmb_dtype dtype, int L, double J, int pbc, cmb_op** out);
/* host-only: the exchange plan of rank `rank` of `nranks` for the matrix-free Heisenberg chain.  Entry k of the
 * outputs (capacity >= 8) describes remote bond k: kind (2 straddle, 3 rank-rank, 4 periodic wrap), partner rank,
 * whether a slab travels, its offset in this rank's receive buffer and its length, both in units of the local slab
 * length / 2.  Returns the number of remote bonds (0 for a single rank), negative on error.  No GPU needed. */
int cmb_heisenberg_plan(int L, int pbc, int nranks, int rank, int* kind, int* partner, int* needed, int* offset_half,
                        int* length_half);
/* legacy host callback with the reference's signature plus a user pointer: out = A*in on LOCAL host
 * slabs (single-rank contexts only).  Costs one D2H + one H2D of an n-vector per Krylov step. */
typedef void (*cmb_matmul_fn)(const void* in, void* out, void* user);
int cmb_op_callback_create(cmb_ctx* ctx, cmb_dtype dtype, int64_t n, cmb_matmul_fn fn, void* user, cmb_op** out);
/* Device-resident operator algebra (the reference composes operators on the host, vector_map.hpp:38-266):
 * sum_i coefs[i] * ops[i] (coefs: nterms dtype elements) and outer * inner, as operators of their own.  The children
 * stay owned by the caller and must outlive the composition; all must share context, dtype, shape and row range. */
int cmb_op_linear_create(cmb_ctx* ctx, int64_t nterms, cmb_op* const* ops, const void* coefs, cmb_op** out);
int cmb_op_product_create(cmb_ctx* ctx, cmb_op* outer, cmb_op* inner, cmb_op** out);
/* collective on multi-rank contexts when the operator exchanges halos through peer memory: every rank must call it */
int cmb_op_destroy(cmb_op* op);
cmb_ctx* cmb_op_context(const cmb_op* op);
int64_t cmb_op_row_begin(const cmb_op* op);
int64_t cmb_op_rows(const cmb_op* op);   /* local rows   */
int64_t cmb_op_height(const cmb_op* op); /* global height */
int cmb_op_dtype(const cmb_op* op);
double cmb_op_bytes(const cmb_op* op);   /* algorithmic bytes of one local apply */
/* y = A x on host slabs (test / debugging entry point; x and y are local slabs) */
int cmb_op_apply_host(cmb_op* op, const void* x, void* y);

/* ---- Krylov state: the data members of LanczosBase (lanczos.hpp:233-239) / ArnoldiBase
 * (arnoldi.hpp:181-187) kept in HBM --------------------------------------------------------------- */
int cmb_krylov_create(cmb_ctx* ctx, cmb_dtype dtype, int64_t n_global, int64_t row_begin, int64_t row_end,
                      int64_t reserve_cols, cmb_krylov** out);
int cmb_krylov_destroy(cmb_krylov* k);
/* clearLanczosSteps() / clearArnoldiSteps() (lanczos.hpp:277-283): drops the basis, keeps deflation vectors */
int cmb_krylov_clear(cmb_krylov* k);
/* orthogonalizingVectors_ (lanczos.hpp:153): nvec host vectors (column-major, leading dimension ld) */
int cmb_krylov_set_deflation(cmb_krylov* k, int64_t nvec, const void* vecs, int64_t ld);
/* setInitialLanczosvector() / setInitialArnoldivector() (lanczos.hpp:299-323, arnoldi.hpp:245-269):
 * copy the local slab of the start vector, project the deflation vectors out, test the norm against
 * threshold and normalise.  *status is CMB_STEP_OK or CMB_STEP_NOSTART. */
int cmb_krylov_start(cmb_krylov* k, const void* init, double threshold, int* status);
/* starts again from the vector of the last cmb_krylov_start, which the state keeps in HBM: no host-to-device copy.
 * CMB_ERR_INVALID when there was no such call since the state was created. */
int cmb_krylov_restart(cmb_krylov* k, double threshold, int* status);
int64_t cmb_krylov_ncols(const cmb_krylov* k); /* number of Krylov vectors */
int64_t cmb_krylov_rows(const cmb_krylov* k);
/* copy Krylov vector j (local slab) to host */
int cmb_krylov_get_col(cmb_krylov* k, int64_t j, void* out);

/* updateLanczosSteps() (lanczos.hpp:371-457).  First call after cmb_krylov_start: v=(A+shift)u0,
 * alpha0.  Later calls: one Lanczos step; full reorthogonalisation (interval 1) runs as CGS2 in three
 * fused passes over the device-resident basis, other intervals follow the reference's strided subset.
 * alpha/beta: values produced by this call (alpha always when status==OK; beta on every later call). */
int cmb_lanczos_step(cmb_krylov* k, cmb_op* op, double shift, int64_t interval, double threshold,
                     double* alpha, double* beta, int* status);
/* up to nsteps steps enqueued without host synchronisation in between (a device-side flag stops the
 * chain at breakdown).  alpha/beta receive the values of the steps done. */
int cmb_lanczos_run(cmb_krylov* k, cmb_op* op, double shift, int64_t interval, double threshold,
                    int64_t nsteps, double* alpha, double* beta, int64_t* steps_done, int* status);

/* ||A u_k - alpha_k u_k - beta_{k-1} u_{k-1}|| for the newest Lanczos vector u_k: the beta the next step would
 * produce, i.e. the factor of the Ritz residual bounds |beta_next * S(last,i)|.  Does not change the state. */
int cmb_lanczos_residual_norm(cmb_krylov* k, double* out);

/* Thick restart (additive; Wu & Simon 2000).  With Krylov vectors u_0..u_m on the device, replaces them by
 * [V_m coef(:,0..nkeep), u_m]: coef is the m x nkeep (column-major, leading dimension ldc, real) matrix of Ritz
 * coefficient vectors of the projected matrix of u_0..u_{m-1}.  v = (A+shift) u_m and alpha_m are kept, so the next
 * cmb_lanczos_run step continues from u_m (full reorthogonalisation only). */
int cmb_lanczos_thick_restart(cmb_krylov* k, const double* coef, int64_t ldc, int64_t m, int64_t nkeep);

/* Thick restart of the Arnoldi iteration (additive; Krylov-Schur, Stewart 2001).  With Arnoldi vectors q_0..q_{m-1} on
 * the device and the residual vector of the last step pending, replaces the basis by Q coef: coef is the m x nkeep
 * (column-major, leading dimension ldc, elements of the basis dtype) orthonormal basis of the subspace of the projected
 * matrix to keep.  The residual vector and its norm stay, so the next cmb_arnoldi_run step continues from it; the caller
 * keeps the projected matrix (coef^H H coef in the leading block, residue * coef(m-1, :) as row nkeep). */
int cmb_arnoldi_thick_restart(cmb_krylov* k, const void* coef, int64_t ldc, int64_t m, int64_t nkeep);

/* updateArnoldiSteps() (arnoldi.hpp:312-392).  hcol receives h(0..ncols-1, ncols-1) of the new column
 * (dtype elements); *residue the new residual norm.  shift points at one dtype element. */
int cmb_arnoldi_step(cmb_krylov* k, cmb_op* op, const void* shift, double threshold, void* hcol,
                     double* residue, int* status);

/* up to nsteps Arnoldi steps enqueued without host synchronisation in between (the device stops the chain when a
 * residue <= threshold appears).  Column j of hcols (leading dimension ldh, dtype elements) receives
 * h(0..k_j, k_j) of the j-th new column, residues[j] its residual norm. */
int cmb_arnoldi_run(cmb_krylov* k, cmb_op* op, const void* shift, double threshold, int64_t nsteps, void* hcols,
                    int64_t ldh, double* residues, int64_t* steps_done, int* status);

/* Ritz-vector assembly (lanczos.hpp:797-817, arnoldi.hpp:841-865): X(:,e) = sum_m coef(m,e) * basis_m,
 * normalised, multiplied by the conjugate phase of its first non-zero element.  coef is column-major
 * ncoef x nev with leading dimension ldc, of coef_dtype; X is written to host memory (local slab,
 * leading dimension ldx) in coef_dtype (a real basis with complex coefficients gives complex vectors). */
int cmb_krylov_ritz_vectors(cmb_krylov* k, cmb_dtype coef_dtype, const void* coef, int64_t ldc,
                            int64_t ncoef, int64_t nev, void* x_host, int64_t ldx);

/* g = V^H x over the Krylov vectors (x: local host slab; g: ncols dtype elements) and out = sum_m coef_m u_m over the
 * first ncoef Krylov vectors (plain combination: no normalisation, no phase).  Together they evaluate functions of
 * the operator in the Krylov space, e.g. LanczosExponentialSolver::solveWithLanczos (lanczos.hpp:1061-1075), with
 * two passes over the basis instead of one host axpy per (Ritz vector, basis vector) pair. */
int cmb_krylov_project(cmb_krylov* k, const void* x_host, void* g_host);
int cmb_krylov_combine(cmb_krylov* k, const void* coef, int64_t ncoef, void* out_host);

/* algorithmic bytes moved by the Krylov steps so far: sum of B_op + (3c+7) n s */
double cmb_krylov_bytes(const cmb_krylov* k);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CMPT_B200_H_ */
