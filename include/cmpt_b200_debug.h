/* cmpt_b200_debug.h — diagnostic and test entry points of libcmpt_b200.so.
 *
 * Not part of the boundary a reference user binds (that is cmpt_b200.h / cmpt_b200_solver.h): these functions exist
 * for the test-suite and for measurements.
 *
 * Virtual ranks: P row-partitioned contexts on ONE device inside ONE process, so that everything that only runs with
 * more than one rank (peer-memory mailboxes of the Gram-Schmidt passes, halo push fused into the SpMV, Pythagorean
 * beta, slab exchange of the matrix-free Heisenberg apply) can be exercised on a box with a single GPU.  Each virtual
 * rank owns a disjoint share of the SMs (CUDA green context) and must be driven by its own host thread; the host-side
 * collectives of cmb_ctx_create_dist contexts (NCCL) become rendezvous of those threads.  The kernels, flags and
 * sequence numbers are exactly those of real ranks.  Set CUDA_DEVICE_MAX_CONNECTIONS=32 before the first CUDA call so
 * that the streams of different virtual ranks do not share a hardware queue.
 */
#ifndef CMPT_B200_DEBUG_H_
#define CMPT_B200_DEBUG_H_

#include "cmpt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct cmb_vgroup cmb_vgroup;
int cmb_vgroup_create(int device, int nranks, cmb_vgroup** out);
int cmb_vgroup_destroy(cmb_vgroup* group);
/* green_contexts: 1 when every virtual rank got its own SM partition, 0 when the ranks share the SMs (grids are then
 * sized for 1/P of the device each) */
int cmb_vgroup_info(const cmb_vgroup* group, int* nranks, int* green_contexts, int* sms_per_rank);
/* collective over the group's ranks: call it from nranks host threads */
int cmb_ctx_create_virtual(cmb_vgroup* group, int rank, cmb_ctx** out);

/* Bound of the in-kernel waits for a peer rank (mailbox and halo flags), default 30 s (CMPT_B200_SPIN_TIMEOUT_S).  A
 * timed-out wait halts the step chain, is reported as CMB_ERR_NCCL by the call that launched it and leaves the
 * context unusable (the ranks' sequence numbers no longer agree). */
int cmb_ctx_set_spin_timeout(cmb_ctx* ctx, double seconds);

/* the row-partitioned Heisenberg operator with nranks virtual ranks on one GPU, sequentially on one stream (same
 * kernels and packing as real ranks, the NVLink exchange replaced by device copies); x, y: full 2^L host vectors */
int cmb_debug_heisenberg_virtual(cmb_ctx* ctx, cmb_dtype dtype, int L, double J, int pbc, int nranks, const void* x,
                                 void* y);
/* storage of a CSR operator after the conversion to SELL-32: entries of the matrix, entries stored (with padding), and
 * whether the rows were sorted by length inside 1024-row windows (SELL-32-1024) */
int cmb_debug_op_sell_stats(cmb_op* op, long long* nnz, long long* padded, int* sorted);
/* number of halo exchanges the CSR shard of this rank has completed (-1 when it has no peer-memory halo) */
int cmb_debug_op_exchange_count(cmb_op* op, long long* count);
/* mean device time of one Gram-Schmidt pass (mode 0 DOT, 1 UPDATE_DOT, 2 UPDATE_NORM) over the first ncols columns
 * of the basis, `reps` back-to-back launches timed with CUDA events */
int cmb_debug_cgs_pass(cmb_krylov* k, int mode, int ncols, int reps, double* ms_per_launch);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CMPT_B200_DEBUG_H_ */
