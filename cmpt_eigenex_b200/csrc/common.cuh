// common.cuh — shared internals of libcmpt_b200.so (context, error plumbing, TMA / mbarrier wrappers).
#pragma once

#include <cuda.h>
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <complex>
#include <map>
#include <string>
#include <vector>

#include "cmpt_b200.h"

namespace cmb {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define CMB_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::cmb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CMB_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define CMB_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != CMB_OK) return _r; \
  } while (0)

#define CMB_REQUIRE(cond, msg)                                         \
  do {                                                                 \
    if (!(cond)) {                                                     \
      ::cmb::set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond); \
      return CMB_ERR_INVALID;                                          \
    }                                                                  \
  } while (0)

// ---- NCCL, loaded at run time (libnccl.so.2) so that single-GPU use needs no NCCL at all --------
struct NcclId {
  char internal[128];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId /* ncclUniqueId, by value */, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
int nccl_load(NcclApi** api);  // CMB_OK or CMB_ERR_NCCL
enum { kNcclFloat64 = 8, kNcclInt8 = 0, kNcclInt32 = 2, kNcclInt64 = 4, kNcclUint64 = 5, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };

struct ProfEntry {
  double ms = 0.0;
  uint64_t launches = 0;
};

// ---- peer-memory mailboxes: the fused "reduce over NVLink" of the Gram-Schmidt coefficients -------------
// Every rank owns a mailbox in its HBM, mapped into all peers through CUDA IPC.  The CTA that finishes a
// Gram-Schmidt pass last pushes this rank's partial coefficient vector straight into every peer's mailbox
// with NVLink stores and then publishes a sequence number; the next kernel waits for the P sequence numbers
// in its own mailbox and sums the P partials in rank order (identical bits on every rank).  No separate
// collective kernel runs between the passes.
constexpr int kMaxPeers = 16;
constexpr int kMailSlots = 4;     // ring of slots, indexed by sequence number
constexpr int kMailStride = 272;  // doubles per (slot, sender)
struct MailPush {                 // producer side (kernel argument)
  int P = 1, rank = 0;
  unsigned long long seq = 0;
  double* data[kMaxPeers];                // slot base in every rank's mailbox (own rank included)
  unsigned long long* flag[kMaxPeers];    // flag array of that slot in every rank's mailbox
};
struct MailPull {                 // consumer side (kernel argument)
  int P = 1;
  unsigned long long seq = 0;
  const double* data = nullptr;           // slot base in the own mailbox: [P][kMailStride]
  const unsigned long long* flag = nullptr;
  double* writeback = nullptr;            // optional plain copy of the reduced values (for the host)
  int* error = nullptr;                   // set to 1 when a peer never arrives (bounded spin)
  int* halt = nullptr;                    // sticky halt flag of the chain: raised together with *error
  long long timeout = 1ll << 36;          // spin bound in SM clocks (cmb_ctx::spin_timeout)
};

// Halo exchange through peer memory (halo.cu): the owner of the rows stores the values a peer needs straight into
// that peer's receive buffer over NVLink and then raises a per-sender flag there; the consumer (the SpMV kernel)
// waits on its own flags.  Receive buffers are double-buffered by the parity of the exchange number.
struct HaloPush {                       // producer side (kernel argument)
  int P = 1, rank = 0;
  long long send_off[kMaxPeers + 1];    // slice of the send-index list that goes to each peer
  double* dst[kMaxPeers];               // where this rank's block starts in the peer's buffer 0
  long long stride[kMaxPeers];          // distance (doubles) between the peer's buffers 0 and 1
  unsigned long long* flag[kMaxPeers];  // this rank's flag in the peer's memory
  unsigned long long* xseq = nullptr;   // local count of executed exchanges (advanced by the consuming kernel)
  unsigned* ticket = nullptr;
  const int* idx = nullptr;             // local rows to send, grouped by peer (send_off)
  int npush = 1;                        // CTAs of the consuming kernel that take part in the push
  int all_push = 0;                     // every CTA pushes a share and then goes on to its rows (few interior slices)
};
struct HaloPull {                       // consumer side (kernel argument)
  const double* base = nullptr;         // receive buffer 0 (or the NCCL receive buffer when flag == nullptr)
  long long stride = 0;
  const unsigned long long* flag = nullptr;  // [P] per-sender flags in local memory; nullptr: nothing to wait for
  const unsigned long long* xseq = nullptr;
  int P = 1, rank = 0;
  int* error = nullptr;
  int* halt = nullptr;                  // sticky halt flag of the chain: raised together with *error
  long long timeout = 1ll << 36;        // spin bound in SM clocks (cmb_ctx::spin_timeout)
};

struct VGroup;  // virtual ranks on one device (vgroup.cu)

}  // namespace cmb

// ---- the context ---------------------------------------------------------------------------------
struct cmb_ctx {
  int device = 0;
  int rank = 0;
  int nranks = 1;
  int num_sms = 148;
  cudaStream_t stream = nullptr;  // every kernel of the library runs here
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  uint64_t launches = 0;
  void* nccl_comm = nullptr;
  cmb::NcclApi* nccl = nullptr;
  // cross-CTA reduction workspace: partial sums [kMaxGrid][kPartialStride] and a ticket counter
  double* d_partial = nullptr;
  unsigned* d_ticket = nullptr;
  // peer-memory mailboxes (multi-rank contexts; see cmb::MailPush)
  bool mail_ok = false;
  double* mail_data[cmb::kMaxPeers] = {};              // [rank]: base of that rank's mailbox data (mapped here)
  unsigned long long* mail_flag[cmb::kMaxPeers] = {};  // [rank]: base of that rank's flag array
  unsigned long long mail_seq = 0;                     // sequence number of the last push (same on all ranks)
  int* d_mail_error = nullptr;
  // Bound of the in-kernel waits for a peer (mailbox, halo flags) in SM clocks: ~30 s by default, CMPT_B200_SPIN_TIMEOUT_S
  // or cmb_ctx_set_spin_timeout() change it.  A kernel replayed by a profiler or stopped in a debugger can exceed any
  // bound: multi-rank runs are not compatible with replaying tools.
  long long spin_timeout = 60000000000ll;
  double norm_guard = 1e-8;        // guard ratio of the Pythagorean norm (CgsPass::norm_guard; CMPT_B200_NORM_GUARD)
  bool dead = false;               // a peer wait timed out: the sequence numbers of the ranks no longer agree
  cudaMemPool_t mempool = nullptr; // virtual ranks: a pool of their own (see pool_alloc)
  std::vector<void*> graveyard, host_graveyard;  // virtual ranks: frees deferred to cmb_ctx_destroy (see dfree)
  cmb::VGroup* vgroup = nullptr;   // virtual rank (vgroup.cu): collectives are host-thread rendezvous, not NCCL
  // L2 flush buffer
  void* d_flush = nullptr;
  size_t flush_bytes = 0;
  // per-family profiling
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  struct Pending {
    std::string family;
    cudaEvent_t a, b;
  };
  std::vector<Pending> pending;
  std::map<std::string, cmb::ProfEntry> prof;
};

namespace cmb {

constexpr int kMaxGrid = 148 * 8;      // upper bound on persistent grids that use d_partial
constexpr int kPartialStride = 2 * 136; // doubles per CTA row of d_partial (complex: re,im per column)

// RAII-ish launch bracket: counts the launch and, when profiling, records events around it.
struct LaunchScope {
  cmb_ctx* ctx;
  cudaEvent_t a = nullptr, b = nullptr;
  const char* family;
  LaunchScope(cmb_ctx* c, const char* fam);
  ~LaunchScope();
};
int resolve_profile(cmb_ctx* ctx);
#ifdef __CUDACC__
// Launch with programmatic stream serialization (see grid_dependency_wait) when `overlap` is set, normally otherwise.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(bool overlap, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(grid));
  cfg.blockDim = dim3(unsigned(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = overlap ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
// Policy: which launches overlap with their predecessor.  Measured on B200 (profiles/r2/pdl_and_sell_sort.md): on the
// Gram-Schmidt passes the overlap is worth +14 % when the kernels last ~10 us, +1.4 % at ~100 us and nothing at 1 ms; on
// the operator kernels it costs 1-2 % (their early CTAs start ahead of the rest and unbalance the static slice split).
// Row-partitioned over 8 GPUs (cfg 2) the run with the overlap was 2 % slower than without (2 019 against 2 064 it/s).
// So: single-rank passes always, operator kernels and multi-rank runs only on request.  CMPT_B200_PDL=0 turns it off,
// CMPT_B200_PDL_APPLY=1 adds the operator kernels, CMPT_B200_PDL_RANKS=1 the row-partitioned runs.
inline bool pdl_wanted(bool is_apply, int nranks) {
  struct Policy {
    bool on = true, apply = false, ranks = false;
    Policy() {
      if (const char* e = getenv("CMPT_B200_PDL")) on = atoi(e) != 0;
      if (const char* e = getenv("CMPT_B200_PDL_APPLY")) apply = atoi(e) != 0;
      if (const char* e = getenv("CMPT_B200_PDL_RANKS")) ranks = atoi(e) != 0;
    }
  };
  static const Policy p;
  return p.on && (!is_apply || p.apply) && (nranks == 1 || p.ranks);
}
#endif

int allreduce_sum_f64(cmb_ctx* ctx, double* dev_ptr, size_t count);  // no-op when nranks == 1
int allreduce_min_u64(cmb_ctx* ctx, unsigned long long* dev_ptr, size_t count);
// mailbox bookkeeping (host): the next push gets a fresh sequence number; a pull names the push it consumes
MailPush mail_next_push(cmb_ctx* ctx);
MailPull mail_pull_of(cmb_ctx* ctx, unsigned long long seq, double* writeback);
// collective CUDA-IPC mapping of one cudaMalloc'ed buffer per rank (all-or-nothing; ctx.cu)
bool ipc_share(cmb_ctx* ctx, void* base, void** mapped);
void ipc_unshare(cmb_ctx* ctx, void** mapped);
// all ranks wait for each other (stream synchronised first); no-op on a single rank
int rank_barrier(cmb_ctx* ctx);
// collective: returns rc when rc is an error, an error when another rank reported one, CMB_OK when nobody did
int agree_status(cmb_ctx* ctx, int rc, const char* what);
// all-to-all of int32 device lists: the piece [send_off[q], send_off[q+1]) of d_send goes to rank q, the piece rank q
// destined to this rank lands at d_recv + recv_off[q]
int alltoallv_i32(cmb_ctx* ctx, const int32_t* d_send, const int64_t* send_off, int32_t* d_recv, const int64_t* recv_off);
// Reports (and clears) a timed-out peer wait of the kernels launched so far; the stream must be synchronised.
int check_peer_wait(cmb_ctx* ctx);
// virtual ranks (vgroup.cu)
int vgroup_allreduce_f64(cmb_ctx* c, double* p, size_t count);
int vgroup_allreduce_min_u64(cmb_ctx* c, unsigned long long* p, size_t count);
bool vgroup_share(cmb_ctx* c, void* base, void** mapped);
int vgroup_barrier(cmb_ctx* c);
int vgroup_alltoallv_i32(cmb_ctx* c, const int32_t* d_send, const int64_t* send_off, int32_t* d_recv, const int64_t* recv_off);
int vgroup_attach(VGroup* g, cmb_ctx* c, int rank);
void vgroup_detach(VGroup* g);
int vgroup_size(const VGroup* g);
int vgroup_device(const VGroup* g);

// driver entry point for tensor-map encoding (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// Stream-ordered allocations from the device's default memory pool (release threshold raised to "never" in
// ctx_init_common), so that building and destroying operators of the same size costs no cudaMalloc/cudaFree.
template <class T>
inline int pool_alloc(cmb_ctx* ctx, T** p, size_t bytes) {
  *p = nullptr;
  // Virtual ranks allocate from a pool of their own: in the device's shared default pool the driver may make this
  // rank's stream wait for the stream of the rank that freed the block it reuses (cudaMemPoolReuseAllowInternal-
  // Dependencies), and that stream may be executing a kernel that waits for this rank: dead-lock.
  cudaError_t e = ctx->mempool ? cudaMallocFromPoolAsync(reinterpret_cast<void**>(p), bytes ? bytes : 16, ctx->mempool, ctx->stream)
                               : cudaMallocAsync(reinterpret_cast<void**>(p), bytes ? bytes : 16, ctx->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("out of device memory (%zu bytes): %s", bytes, cudaGetErrorString(e));
    return CMB_ERR_NOMEM;
  }
  return CMB_OK;
}
inline void pool_free(cmb_ctx* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}

// cudaFree / cudaFreeHost synchronise the whole device.  Between real ranks (one process per GPU) that is harmless; a
// virtual rank that frees something while a peer's kernel spins on this rank's NEXT launch would dead-lock, so virtual
// ranks park the pointers until cmb_ctx_destroy (all ranks idle there).  Test-only mode: the growth is bounded by the
// test's allocations.
inline void dfree(cmb_ctx* ctx, void* p) {
  if (!p) return;
  if (ctx && ctx->vgroup)
    ctx->graveyard.push_back(p);
  else
    cudaFree(p);
}
inline void hfree(cmb_ctx* ctx, void* p) {
  if (!p) return;
  if (ctx && ctx->vgroup)
    ctx->host_graveyard.push_back(p);
  else
    cudaFreeHost(p);
}
// small device -> host / memset transfers on the context's stream (never on the legacy default stream: under green
// contexts a synchronous copy there waits for the kernels of every virtual rank)
inline int d2h_sync(cmb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CMB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CMB_OK;
}

template <class T>
inline T* dalloc(size_t n) {
  T* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
  return p;
}

#ifdef __CUDACC__
// ---- device-side PTX wrappers: mbarrier + TMA -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 2D tiled TMA load: coordinates {c0 (innermost: row), c1 (column)}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// 3D tiled TMA load: coordinates {c0 (innermost), c1, c2}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// 1D bulk copy global -> shared (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// The same two copies with an L2 eviction policy (createpolicy) for the lines they bring in
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], "
      "[%5], %6;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// Programmatic dependent launch.  A kernel launched with launch_pdl() may start while the kernel before it on the
// stream is still draining: everything it does before grid_dependency_wait() must not depend on that kernel's output
// (or on anything older kernels wrote that the previous one could still be writing).  Every CTA of such a kernel calls
// grid_dependency_wait() before it reads or writes global memory other than read-only operator/basis data, and before
// it exits.  grid_launch_dependents() lets the scheduler start the next kernel's CTAs as SMs free up; it is issued
// after the kernel's own wait, so that a dependent's early work never overlaps the kernel two places back.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif  // __CUDACC__

}  // namespace cmb
