// ops.cu — operator application kernels: y = (A + shift) x fused with the step's normalisation and the
// Lanczos alpha dot.  These replace the user `matmul` callback and the three lines around it
// (lanczos.hpp:439-448, arnoldi.hpp:365-372): u = w/beta ; v = A u ; v += shift u ; alpha = <u|v>.
//   - CSR input converted on the device to SELL-32 (32-row slices, column-major inside a slice):
//     thread-per-row, fully coalesced value/index streams, x gathered through L1/L2; on request the rows of every
//     1024-row window are sorted by length first (SELL-32-1024: almost no padding, but scattered gathers);
//   - dense row-major GEMV (cfg 1), warp per row;
//   - matrix-free spin-1/2 Heisenberg chain (cfg 5): heisenberg.cu;
//   - legacy host callback (reference signature), staged through pinned host memory.
// All are HBM-bound; algorithmic bytes per apply are recorded in cmb_op::bytes.
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "cmpt_b200_debug.h"
#include "device_utils.cuh"
#include "halo.cuh"
#include "op.cuh"

namespace cmb {

// ======================================================================================================
// SELL-32
// ======================================================================================================
// DIST: row-partitioned shard with the peer-memory halo fused in (a separate instantiation keeps the single-GPU
// kernel free of that bookkeeping).  32 registers either way: the gathers are latency bound, occupancy is what counts.
//
// DIST kernels: the first hpush.npush CTAs only push this rank's boundary values into the peers' receive buffers
// (halo_push_part).  The others walk the slices in `order`: first the n_interior slices that read no remote column,
// then — after waiting for the peers' flags — the boundary slices.  The CTA that leaves last counts the exchange.
struct SellDist {
  unsigned long long seq;
  int npush;
  bool fused;
};

template <int ES>
__device__ __forceinline__ bool sell_dist_prologue(const HaloPush& hpush, const HaloPull& hp, const double* w,
                                                   double* partial, unsigned* ticket, double* alpha_slot, SellDist& d) {
  d.fused = hp.flag != nullptr;
  d.seq = 0;
  d.npush = 0;
  if (!d.fused) return true;
  d.seq = *reinterpret_cast<const volatile unsigned long long*>(hp.xseq) + 1ull;
  if (hpush.all_push) {  // nothing to overlap with: spread the push over the whole grid, then everybody computes
    halo_push_part<ES>(hpush, w, d.seq);
    return true;
  }
  d.npush = hpush.npush;
  if (int(blockIdx.x) < d.npush) {
    halo_push_part<ES>(hpush, w, d.seq);
    if (grid_sum_finalize<ES>(0.0, 0.0, partial, ticket, alpha_slot) && threadIdx.x == 0) *hpush.xseq = d.seq;
    return false;  // pusher CTAs are done
  }
  return true;
}

// PERM: lane l of slice s works on row perm[32 s + l] (rows sorted by length inside 1024-row windows; -1 = no row).  The
// vectors keep their natural order: the lanes of a warp read w and write u, v at 32 places of one 8 KB window.
template <bool CPLX, bool DIST, bool PERM>
__global__ void __launch_bounds__(256, 8)
spmv_sell_kernel(const long long* __restrict__ slice_ptr, const int* __restrict__ col, const double* __restrict__ val,
                 const int* __restrict__ perm, long long nrows, int nslices, const double* __restrict__ w,
                 HaloPush hpush, HaloPull hp,
                 const int* __restrict__ order, int n_interior, double* __restrict__ ucol, double* __restrict__ v,
                 double shr, double shi, StepScalars sc, double* partial, unsigned* ticket) {
  grid_dependency_wait();  // w, the step scalars and the halt flag come from the kernels before this one
  grid_launch_dependents();
  double inv;
  if (!step_prologue(sc, inv)) return;
  SellDist dist{0ull, 0, false};
  if (DIST && !sell_dist_prologue<CPLX ? 2 : 1>(hpush, hp, w, partial, ticket, sc.alpha_slot, dist)) return;
  const int lane = threadIdx.x & 31;
  const int gwarp = (int(blockIdx.x) - dist.npush) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = (int(gridDim.x) - dist.npush) * (blockDim.x >> 5);
  double d0 = 0.0, d1 = 0.0;
  auto body = [&](long long slice, const double* halo) {
    const long long base = slice_ptr[slice];
    const int width = int((slice_ptr[slice + 1] - base) >> 5);
    long long r = slice * 32 + lane;
    if (PERM) {
      r = perm[r];
      if (r < 0) r = nrows;
    }
    const int* cp = col + base + lane;
    if (CPLX) {
      const double2* vp = reinterpret_cast<const double2*>(val) + base + lane;
      const double2* wz = reinterpret_cast<const double2*>(w);
      const double2* hz = reinterpret_cast<const double2*>(halo);
      double ar = 0.0, ai = 0.0;
#pragma unroll 4
      for (int k = 0; k < width; ++k) {
        const int c = __ldg(cp + k * 32);
        const double2 a = __ldg(vp + k * 32);
        const double2 xv = (c < nrows) ? wz[c] : hz[c - nrows];
        ar = fma(a.x, xv.x, ar);
        ar = fma(-a.y, xv.y, ar);
        ai = fma(a.x, xv.y, ai);
        ai = fma(a.y, xv.x, ai);
      }
      if (r < nrows) {
        const double2 wi = wz[r];
        const double ur = wi.x * inv, ui = wi.y * inv;
        const double yr = ar * inv + (shr * ur - shi * ui);
        const double yi = ai * inv + (shr * ui + shi * ur);
        reinterpret_cast<double2*>(ucol)[r] = make_double2(ur, ui);
        reinterpret_cast<double2*>(v)[r] = make_double2(yr, yi);
        d0 += ur * yr + ui * yi;  // conj(u) * y
        d1 += ur * yi - ui * yr;
      }
    } else {
      const double* vp = val + base + lane;
      double acc = 0.0;
#pragma unroll 4
      for (int k = 0; k < width; ++k) {
        const int c = __ldg(cp + k * 32);
        const double a = __ldg(vp + k * 32);
        const double xv = (c < nrows) ? w[c] : halo[c - nrows];
        acc = fma(a, xv, acc);
      }
      if (r < nrows) {
        const double ui = w[r] * inv;
        const double y = acc * inv + shr * ui;
        ucol[r] = ui;
        v[r] = y;
        d0 = fma(ui, y, d0);
      }
    }
  };
  if (!DIST) {
    for (int gi = gwarp; gi < nslices; gi += nwarps) body(gi, nullptr);
  } else if (!dist.fused) {  // NCCL fallback: the halo values are already in hp.base
    for (int gi = gwarp; gi < nslices; gi += nwarps) body(gi, hp.base);
  } else {
    int gi = gwarp;
    for (; gi < n_interior; gi += nwarps) body(order[gi], nullptr);  // no remote column in these slices
    halo_wait_cta(hp, dist.seq);
    const double* halo = hp.base + (dist.seq & 1ull) * hp.stride;
    for (; gi < nslices; gi += nwarps) body(order[gi], halo);
  }
  const bool last = grid_sum_finalize<CPLX ? 2 : 1>(d0, d1, partial, ticket, sc.alpha_slot);
  if (DIST && dist.fused && last && threadIdx.x == 0) *hpush.xseq = dist.seq;  // this exchange is consumed
}

// Uniform-width variant (every slice has exactly W columns: stencils such as the 5-point Laplacian or the 7-point
// convection-diffusion operator).  W is a compile-time constant, so all W index loads, W value loads and then all W
// gathers of a row are issued back to back (one dependent phase instead of a runtime loop) and no slice pointer is
// read.  Real scalars only; everything else goes through the generic kernel.
template <int W, bool DIST>
__global__ void __launch_bounds__(256, 8)
spmv_sell_uniform_kernel(const int* __restrict__ col, const double* __restrict__ val, long long nrows, int nslices,
                         const double* __restrict__ w, HaloPush hpush, HaloPull hp, const int* __restrict__ order,
                         int n_interior, double* __restrict__ ucol, double* __restrict__ v, double shr,
                         StepScalars sc, double* partial, unsigned* ticket) {
  grid_dependency_wait();  // w, the step scalars and the halt flag come from the kernels before this one
  grid_launch_dependents();
  double inv;
  if (!step_prologue(sc, inv)) return;
  SellDist dist{0ull, 0, false};
  if (DIST && !sell_dist_prologue<1>(hpush, hp, w, partial, ticket, sc.alpha_slot, dist)) return;
  const int lane = threadIdx.x & 31;
  const int gwarp = (int(blockIdx.x) - dist.npush) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = (int(gridDim.x) - dist.npush) * (blockDim.x >> 5);
  double d0 = 0.0;
  auto body = [&](long long slice, const double* halo) {
    const long long base = slice * (W * 32) + lane;
    const long long r = slice * 32 + lane;
    int c[W];
    double a[W], xv[W];
#pragma unroll
    for (int k = 0; k < W; ++k) c[k] = __ldg(col + base + k * 32);
#pragma unroll
    for (int k = 0; k < W; ++k) a[k] = __ldg(val + base + k * 32);
#pragma unroll
    for (int k = 0; k < W; ++k) xv[k] = (!DIST || c[k] < nrows) ? w[c[k]] : halo[c[k] - nrows];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < W; ++k) acc = fma(a[k], xv[k], acc);
    if (r < nrows) {
      const double ui = w[r] * inv;
      const double y = acc * inv + shr * ui;
      ucol[r] = ui;
      v[r] = y;
      d0 = fma(ui, y, d0);
    }
  };
  if (!DIST) {
    for (int gi = gwarp; gi < nslices; gi += nwarps) body(gi, nullptr);
  } else if (!dist.fused) {
    for (int gi = gwarp; gi < nslices; gi += nwarps) body(gi, hp.base);
  } else {
    int gi = gwarp;
    for (; gi < n_interior; gi += nwarps) body(order[gi], nullptr);
    halo_wait_cta(hp, dist.seq);
    const double* halo = hp.base + (dist.seq & 1ull) * hp.stride;
    for (; gi < nslices; gi += nwarps) body(order[gi], halo);
  }
  const bool last = grid_sum_finalize<1>(d0, 0.0, partial, ticket, sc.alpha_slot);
  if (DIST && dist.fused && last && threadIdx.x == 0) *hpush.xseq = dist.seq;
}

// flag[s] = 1 when slice s references a remote (halo) column, i.e. an index >= nrows
__global__ void sell_halo_flag_kernel(const long long* __restrict__ slice_ptr, const int* __restrict__ col, long long nrows,
                                      long long nslices, unsigned char* __restrict__ flag) {
  const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gwarp >= nslices) return;
  int any = 0;
  for (long long i = slice_ptr[gwarp] + lane; i < slice_ptr[gwarp + 1]; i += 32) any |= (col[i] >= nrows);
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) flag[gwarp] = (unsigned char)any;
}

// bad[0] is raised when the CSR arrays are inconsistent (decreasing rowptr, row longer than 2^31, column index outside
// [0, ncols)): the build then fails with CMB_ERR_INVALID instead of leaving an operator that reads out of bounds.
__global__ void sell_width_kernel(const long long* __restrict__ rowptr, const int* __restrict__ perm, long long nrows,
                                  long long nslices, int* __restrict__ width, int* __restrict__ bad) {
  const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gwarp >= nslices) return;
  long long r = gwarp * 32 + lane;
  if (perm) r = perm[r] < 0 ? nrows : perm[r];
  const long long len64 = (r < nrows) ? rowptr[r + 1] - rowptr[r] : 0;
  if (len64 < 0 || len64 > 0x7fffffffll) *bad = 1;
  int len = (len64 < 0 || len64 > 0x7fffffffll) ? 0 : int(len64);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) width[gwarp] = len;
}

// SELL-32-1024: one CTA sorts the rows of one 1024-row window by length, longest first, ties in row order (a stable
// rank: every thread counts the rows that go before its own).  perm[position] = row, -1 behind the last row.
constexpr int kSellSigma = 1024;
__global__ void __launch_bounds__(kSellSigma)
sell_sort_window_kernel(const long long* __restrict__ rowptr, long long nrows, int* __restrict__ perm) {
  __shared__ int s_len[kSellSigma];
  const long long r = (long long)blockIdx.x * kSellSigma + threadIdx.x;
  long long len64 = (r < nrows) ? rowptr[r + 1] - rowptr[r] : -1;
  if (len64 > 0x7fffffffll) len64 = 0x7fffffffll;  // reported as an error by sell_width_kernel
  const int len = len64 < -1 ? -1 : int(len64);
  s_len[threadIdx.x] = len;
  __syncthreads();
  int rank = 0;
  for (int j = 0; j < kSellSigma; ++j) {
    const int lj = s_len[j];
    rank += (lj > len || (lj == len && j < int(threadIdx.x))) ? 1 : 0;
  }
  perm[(long long)blockIdx.x * kSellSigma + rank] = (r < nrows) ? int(r) : -1;
}

template <int ES>
__global__ void sell_fill_kernel(const long long* __restrict__ rowptr, const int* __restrict__ col,
                                 const double* __restrict__ val, const int* __restrict__ perm, long long nrows,
                                 long long nslices, const long long* __restrict__ slice_ptr, int* __restrict__ scol,
                                 double* __restrict__ sval, long long ncols, int* __restrict__ bad) {
  const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gwarp >= nslices) return;
  long long r = gwarp * 32 + lane;
  if (perm) r = perm[r] < 0 ? nrows : perm[r];
  const long long base = slice_ptr[gwarp];
  const int width = int((slice_ptr[gwarp + 1] - base) >> 5);
  long long p0 = 0;
  int len = 0;
  if (r < nrows) {
    p0 = rowptr[r];
    len = int(rowptr[r + 1] - p0);
  }
  const int self = int(r < nrows ? r : 0);
  for (int k = 0; k < width; ++k) {
    const long long dst = base + (long long)k * 32 + lane;
    if (k < len) {
      const int cc = col[p0 + k];
      if (cc < 0 || cc >= ncols) *bad = 2;
      scol[dst] = (cc < 0 || cc >= ncols) ? self : cc;
#pragma unroll
      for (int e = 0; e < ES; ++e) sval[dst * ES + e] = val[(p0 + k) * ES + e];
    } else {
      scol[dst] = self;  // padding: zero value, harmless in-range column
#pragma unroll
      for (int e = 0; e < ES; ++e) sval[dst * ES + e] = 0.0;
    }
  }
}

struct SellOp : cmb_op {
  long long nslices = 0;
  long long* d_slice_ptr = nullptr;
  int* d_col = nullptr;
  double* d_val = nullptr;
  HaloExchange* halo = nullptr;  // row-partitioned shards only
  int* d_perm = nullptr;         // SELL-32-1024: row handled by each lane of each slice (nullptr: natural order)
  int* d_order = nullptr;        // peer-memory halo: slices that touch no remote column first
  long long n_interior = 0;
  int resident_ctas = 0;         // CTAs of the SpMV kernel that fit on the GPU at once (queried on first use)
  long long padded_nnz = 0, nnz = 0;
  int uniform_width = 0;  // > 0: every slice has this width (fast path for real scalars)
  ~SellOp() override {
    pool_free(ctx, d_slice_ptr);
    pool_free(ctx, d_col);
    pool_free(ctx, d_val);
    pool_free(ctx, d_order);
    pool_free(ctx, d_perm);
    delete halo;
  }
  int apply(const double* w, double* ucol, double* v, double shr, double shi, const StepScalars& sc) override {
    long long blocks = (nslices + 7) / 8;
    int grid = int(std::min<long long>(blocks, (long long)ctx->num_sms * 8));
    if (grid < 1) grid = 1;
    HaloPull d_halo;
    HaloPush d_push;
    if (halo) {
      // NVLink halo exchange of the un-normalised w (1/beta is applied inside the SpMV): with peer memory the SpMV
      // kernel gets extra leading CTAs that push; a pack kernel + ncclSend/ncclRecv group otherwise
      CMB_TRY(halo->exchange(ctx, w, sc.halt));
      d_halo = halo->pull_args();
      d_halo.halt = sc.halt;
      if (halo->p2p) {
        if (2 * n_interior < nslices) {
          // most slices read remote columns, so there is little to hide the exchange behind: all CTAs push first.
          // Every CTA then waits for the peers, so the whole grid must be resident at once (a CTA that could not
          // start before others finish would never push): cap the grid by the measured occupancy.
          if (resident_ctas == 0) {
            int per_sm = 0;
            CMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_sell_kernel<true, true, true>, 256, 0));
            resident_ctas = std::max(1, per_sm) * ctx->num_sms;
          }
          grid = std::min(grid, resident_ctas);
          d_push = halo->fused_push(grid);
          d_push.npush = grid;
          d_push.all_push = 1;
        } else {
          d_push = halo->fused_push(ctx->num_sms);
          grid = std::max(1, std::min(grid, kMaxGrid - d_push.npush)) + d_push.npush;
        }
      }
    }
    LaunchScope ls(ctx, "spmv_sell");
#define CMB_SELL_UNIFORM(W)                                                                                        \
  case W:                                                                                                          \
    if (halo)                                                                                                      \
      CMB_CUDA(launch_pdl(pdl_wanted(true, ctx->nranks), spmv_sell_uniform_kernel<W, true>, grid, 256, 0, ctx->stream, d_col, d_val, n_local,     \
                          int(nslices), w, d_push, d_halo, d_order, int(n_interior), ucol, v, shr, sc,             \
                          ctx->d_partial, ctx->d_ticket + 1));                                                     \
    else                                                                                                           \
      CMB_CUDA(launch_pdl(pdl_wanted(true, ctx->nranks), spmv_sell_uniform_kernel<W, false>, grid, 256, 0, ctx->stream, d_col, d_val, n_local,    \
                          int(nslices), w, d_push, d_halo, d_order, int(n_interior), ucol, v, shr, sc,             \
                          ctx->d_partial, ctx->d_ticket + 1));                                                     \
    return CMB_OK;
    if (!cplx && shi == 0.0) {
      switch (uniform_width) {
        CMB_SELL_UNIFORM(1)
        CMB_SELL_UNIFORM(2)
        CMB_SELL_UNIFORM(3)
        CMB_SELL_UNIFORM(4)
        CMB_SELL_UNIFORM(5)
        CMB_SELL_UNIFORM(6)
        CMB_SELL_UNIFORM(7)
        CMB_SELL_UNIFORM(8)
        CMB_SELL_UNIFORM(9)
        default:
          break;
      }
    }
#undef CMB_SELL_UNIFORM
#define CMB_SELL_GENERIC(C, D)                                                                                    \
  do {                                                                                                            \
    if (d_perm)                                                                                                   \
      CMB_CUDA(launch_pdl(pdl_wanted(true, ctx->nranks), spmv_sell_kernel<C, D, true>, grid, 256, 0, ctx->stream, d_slice_ptr, d_col, d_val,     \
                          d_perm, n_local, int(nslices), w, d_push, d_halo, d_order, int(n_interior), ucol, v,    \
                          shr, shi, sc, ctx->d_partial, ctx->d_ticket + 1));                                      \
    else                                                                                                          \
      CMB_CUDA(launch_pdl(pdl_wanted(true, ctx->nranks), spmv_sell_kernel<C, D, false>, grid, 256, 0, ctx->stream, d_slice_ptr, d_col, d_val,    \
                          d_perm, n_local, int(nslices), w, d_push, d_halo, d_order, int(n_interior), ucol, v,    \
                          shr, shi, sc, ctx->d_partial, ctx->d_ticket + 1));                                      \
  } while (0)
    if (cplx) {
      if (halo)
        CMB_SELL_GENERIC(true, true);
      else
        CMB_SELL_GENERIC(true, false);
    } else {
      if (halo)
        CMB_SELL_GENERIC(false, true);
      else
        CMB_SELL_GENERIC(false, false);
    }
#undef CMB_SELL_GENERIC
    CMB_CUDA(cudaGetLastError());
    return CMB_OK;
  }
};

// The CSR arrays of a shard in HBM (stream-ordered pool allocations), as uploaded from the caller's host arrays.
struct CsrOnDevice {
  cmb_ctx* ctx = nullptr;
  long long* rowptr = nullptr;
  int* col = nullptr;
  double* val = nullptr;
  long long nnz = 0;
  ~CsrOnDevice() {
    if (!ctx) return;
    pool_free(ctx, rowptr);
    pool_free(ctx, col);
    pool_free(ctx, val);
  }
};
static int upload_csr(SellOp* op, const int64_t* rowptr, const int32_t* col, const void* val, CsrOnDevice& d) {
  cmb_ctx* ctx = op->ctx;
  const int es = op->cplx ? 2 : 1;
  const long long n = op->n_local;
  d.ctx = ctx;
  d.nnz = rowptr[n];
  CMB_TRY(pool_alloc(ctx, &d.rowptr, sizeof(long long) * (n + 1)));
  CMB_TRY(pool_alloc(ctx, &d.col, sizeof(int) * std::max<long long>(d.nnz, 1)));
  CMB_TRY(pool_alloc(ctx, &d.val, sizeof(double) * es * std::max<long long>(d.nnz, 1)));
  CMB_CUDA(cudaMemcpyAsync(d.rowptr, rowptr, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
  CMB_CUDA(cudaMemcpyAsync(d.col, col, sizeof(int) * d.nnz, cudaMemcpyHostToDevice, ctx->stream));
  CMB_CUDA(cudaMemcpyAsync(d.val, val, sizeof(double) * es * d.nnz, cudaMemcpyHostToDevice, ctx->stream));
  return CMB_OK;
}

// ncols: number of valid column indices (n_global on a single rank, n_local + halo entries for a shard)
static int build_sell(SellOp* op, const CsrOnDevice& csr, long long ncols) {
  cmb_ctx* ctx = op->ctx;
  const int es = op->cplx ? 2 : 1;
  const long long n = op->n_local;
  const long long nnz = csr.nnz;
  op->nnz = nnz;
  op->nslices = (n + 31) / 32;
  long long* d_rowptr = csr.rowptr;
  int* d_ccol = csr.col;
  double* d_cval = csr.val;
  int* d_width = nullptr;
  int* d_bad = nullptr;
  auto cleanup = [&]() {
    pool_free(ctx, d_width);
    pool_free(ctx, d_bad);
  };
  int rc = [&]() -> int {
    CMB_TRY(pool_alloc(ctx, &d_width, sizeof(int) * std::max<long long>(op->nslices, 1)));
    CMB_TRY(pool_alloc(ctx, &d_bad, sizeof(int)));
    CMB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    const int threads = 256;
    const long long nthreads = op->nslices * 32;
    const int grid = int((nthreads + threads - 1) / threads);
    std::vector<int> width(op->nslices);
    // slice widths for the current row order (op->d_perm), validated; fills `width` and returns the padded size
    auto measure = [&](long long* padded) -> int {
      if (op->nslices > 0) {
        LaunchScope ls(ctx, "sell_build");
        sell_width_kernel<<<grid, threads, 0, ctx->stream>>>(d_rowptr, op->d_perm, n, op->nslices, d_width, d_bad);
      }
      CMB_CUDA(cudaGetLastError());
      int bad = 0;
      CMB_CUDA(cudaMemcpyAsync(width.data(), d_width, sizeof(int) * op->nslices, cudaMemcpyDeviceToHost, ctx->stream));
      CMB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CMB_CUDA(cudaStreamSynchronize(ctx->stream));
      if (bad) {
        set_error("CSR rowptr is not non-decreasing (or a row is longer than 2^31 entries)");
        return CMB_ERR_INVALID;
      }
      *padded = 0;
      for (long long s = 0; s < op->nslices; ++s) *padded += (long long)width[s] * 32;
      return CMB_OK;
    };
    long long padded = 0;
    CMB_TRY(measure(&padded));
    // stencil-like matrices: pad every slice to the maximum width when that costs < 5 % extra entries, which
    // enables the fully unrolled uniform-width kernel
    int maxw = 0;
    for (long long s = 0; s < op->nslices; ++s) maxw = std::max(maxw, width[s]);
    const bool stencil =
        maxw >= 1 && maxw <= 9 && double(maxw) * 32.0 * double(op->nslices) <= 1.05 * double(std::max<long long>(nnz, 1));
    if (stencil) {
      for (long long s = 0; s < op->nslices; ++s) width[s] = maxw;
    } else if (double(padded) > 1.05 * double(std::max<long long>(nnz, 1)) && getenv("CMPT_B200_SELL_SORT") &&
               atoi(getenv("CMPT_B200_SELL_SORT")) != 0) {
      // Irregular rows, on request (CMPT_B200_SELL_SORT=1): sort every 1024-row window by length and keep the order if
      // it saves padding.  Off by default: for the Heisenberg ring (cfg 4) it cuts the stored entries from 1.25 nnz to
      // 1.02 nnz but the SpMV gets 19 % slower (1.23 ms against 1.03 ms at L = 24), because neighbouring lanes no
      // longer gather neighbouring entries of x — that kernel is bound by the gathers, not by the matrix stream.
      const long long nwin = (n + kSellSigma - 1) / kSellSigma;
      CMB_TRY(pool_alloc(ctx, &op->d_perm, sizeof(int) * size_t(std::max<long long>(nwin * kSellSigma, op->nslices * 32))));
      {
        LaunchScope ls(ctx, "sell_build");
        sell_sort_window_kernel<<<unsigned(nwin), kSellSigma, 0, ctx->stream>>>(d_rowptr, n, op->d_perm);
      }
      CMB_CUDA(cudaGetLastError());
      const long long natural = padded;
      std::vector<int> natural_width = width;
      CMB_TRY(measure(&padded));
      if (padded >= natural) {
        pool_free(ctx, op->d_perm);
        op->d_perm = nullptr;
        width = natural_width;
      }
    }
    std::vector<long long> sp(op->nslices + 1);
    sp[0] = 0;
    for (long long s = 0; s < op->nslices; ++s) sp[s + 1] = sp[s] + (long long)width[s] * 32;
    op->padded_nnz = sp[op->nslices];
    op->uniform_width = (op->nslices > 0 && !op->d_perm) ? width[0] : 0;
    for (long long s = 0; s < op->nslices; ++s)
      if (width[s] != op->uniform_width) {
        op->uniform_width = 0;
        break;
      }
    CMB_TRY(pool_alloc(ctx, &op->d_slice_ptr, sizeof(long long) * (op->nslices + 1)));
    CMB_TRY(pool_alloc(ctx, &op->d_col, sizeof(int) * std::max<long long>(op->padded_nnz, 1)));
    CMB_TRY(pool_alloc(ctx, &op->d_val, sizeof(double) * es * std::max<long long>(op->padded_nnz, 1)));
    CMB_CUDA(cudaMemcpyAsync(op->d_slice_ptr, sp.data(), sizeof(long long) * (op->nslices + 1), cudaMemcpyHostToDevice,
                             ctx->stream));
    if (op->nslices > 0) {
      LaunchScope ls(ctx, "sell_build");
      if (es == 2)
        sell_fill_kernel<2><<<grid, threads, 0, ctx->stream>>>(d_rowptr, d_ccol, d_cval, op->d_perm, n, op->nslices,
                                                               op->d_slice_ptr, op->d_col, op->d_val, ncols, d_bad);
      else
        sell_fill_kernel<1><<<grid, threads, 0, ctx->stream>>>(d_rowptr, d_ccol, d_cval, op->d_perm, n, op->nslices,
                                                               op->d_slice_ptr, op->d_col, op->d_val, ncols, d_bad);
    }
    CMB_CUDA(cudaGetLastError());
    int bad = 0;
    CMB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) {
      set_error("CSR column index outside [0, %lld)", ncols);
      return CMB_ERR_INVALID;
    }
    return CMB_OK;
  }();
  cleanup();
  return rc;
}

// Peer-memory halo: permutation of the slices with the ones that read no remote column first, so that the SpMV
// kernel only has to wait for its peers when it reaches the boundary slices.
static int build_slice_order(SellOp* op) {
  cmb_ctx* ctx = op->ctx;
  const long long ns = op->nslices;
  if (ns == 0) return CMB_OK;
  unsigned char* d_flag = nullptr;
  CMB_TRY(pool_alloc(ctx, &d_flag, size_t(ns)));
  const int grid = int((ns * 32 + 255) / 256);
  sell_halo_flag_kernel<<<grid, 256, 0, ctx->stream>>>(op->d_slice_ptr, op->d_col, op->n_local, ns, d_flag);
  std::vector<unsigned char> flag(static_cast<size_t>(ns));
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(flag.data(), d_flag, size_t(ns), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  pool_free(ctx, d_flag);
  if (e != cudaSuccess) {
    set_error("slice ordering failed: %s", cudaGetErrorString(e));
    return CMB_ERR_CUDA;
  }
  std::vector<int> order;
  order.reserve(static_cast<size_t>(ns));
  for (long long s = 0; s < ns; ++s)
    if (!flag[size_t(s)]) order.push_back(int(s));
  op->n_interior = (long long)order.size();
  for (long long s = 0; s < ns; ++s)
    if (flag[size_t(s)]) order.push_back(int(s));
  CMB_TRY(pool_alloc(ctx, &op->d_order, sizeof(int) * size_t(ns)));
  CMB_CUDA(cudaMemcpyAsync(op->d_order, order.data(), sizeof(int) * size_t(ns), cudaMemcpyHostToDevice, ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CMB_OK;
}

// ======================================================================================================
// dense row-major GEMV (single rank)
// ======================================================================================================
// K warps share a row (K = 1 by default; 2, 4, 8 for measurements).  A warp takes a contiguous even-length piece of the
// row, reads it with four independent 16-byte loads per lane in flight, and the first warp of the row adds the K
// partial sums in a fixed order.
template <bool CPLX>
__global__ void __launch_bounds__(256)
dense_apply_kernel(const double* __restrict__ A, long long n, const double* __restrict__ w, double* __restrict__ ucol,
                   double* __restrict__ v, double shr, double shi, StepScalars sc, double* partial, unsigned* ticket,
                   int K) {
  double inv;
  if (!step_prologue(sc, inv)) return;
  __shared__ double s_part[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows_per_cta = 8 / K, sub = warp % K, rloc = warp / K;
  const long long seg = ((n + K - 1) / K + 1) & ~1ll;  // even, so that every piece starts on a 16-byte boundary
  const long long c0 = std::min<long long>(n, sub * seg), c1 = std::min<long long>(n, c0 + seg);
  double d0 = 0.0, d1 = 0.0;
  for (long long rbase = (long long)blockIdx.x * rows_per_cta; rbase < n; rbase += (long long)gridDim.x * rows_per_cta) {
    const long long r = rbase + rloc;
    double ar = 0.0, ai = 0.0;
    if (r < n) {
      if (CPLX) {
        const double2* row = reinterpret_cast<const double2*>(A) + r * n;
        const double2* wz = reinterpret_cast<const double2*>(w);
        double br = 0.0, bi = 0.0;
        long long c = c0 + lane;
        for (; c + 32 < c1; c += 64) {
          const double2 a0 = __ldg(row + c), a1 = __ldg(row + c + 32);
          const double2 x0 = wz[c], x1 = wz[c + 32];
          ar = fma(a0.x, x0.x, ar), ar = fma(-a0.y, x0.y, ar), ai = fma(a0.x, x0.y, ai), ai = fma(a0.y, x0.x, ai);
          br = fma(a1.x, x1.x, br), br = fma(-a1.y, x1.y, br), bi = fma(a1.x, x1.y, bi), bi = fma(a1.y, x1.x, bi);
        }
        if (c < c1) {
          const double2 a0 = __ldg(row + c), x0 = wz[c];
          ar = fma(a0.x, x0.x, ar), ar = fma(-a0.y, x0.y, ar), ai = fma(a0.x, x0.y, ai), ai = fma(a0.y, x0.x, ai);
        }
        ar += br, ai += bi;
      } else if ((n & 1) == 0) {
        const double2* row = reinterpret_cast<const double2*>(A + r * n + c0);
        const double2* wz = reinterpret_cast<const double2*>(w + c0);
        const long long cnt = (c1 - c0) >> 1;
        double a[4] = {0.0, 0.0, 0.0, 0.0};
        for (long long i = lane; i < cnt; i += 128) {
          double2 m[4], x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = i + 32 * u < cnt;
            m[u] = ok ? __ldg(row + i + 32 * u) : make_double2(0.0, 0.0);
            x[u] = ok ? wz[i + 32 * u] : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) a[u] = fma(m[u].y, x[u].y, fma(m[u].x, x[u].x, a[u]));
        }
        ar = (a[0] + a[1]) + (a[2] + a[3]);
      } else {  // odd n: rows are not 16-byte aligned
        const double* row = A + r * n;
        double b = 0.0;
        long long c = c0 + lane;
        for (; c + 32 < c1; c += 64) {
          ar = fma(__ldg(row + c), w[c], ar);
          b = fma(__ldg(row + c + 32), w[c + 32], b);
        }
        if (c < c1) ar = fma(__ldg(row + c), w[c], ar);
        ar += b;
      }
    }
    ar = warp_sum(ar);
    if (CPLX) ai = warp_sum(ai);
    if (K > 1) {
      if (lane == 0) {
        s_part[0][warp] = ar;
        s_part[1][warp] = ai;
      }
      __syncthreads();
      if (sub == 0 && lane == 0) {
        ar = 0.0, ai = 0.0;
        for (int k = 0; k < K; ++k) ar += s_part[0][warp + k], ai += s_part[1][warp + k];
      }
      __syncthreads();
    }
    if (sub == 0 && lane == 0 && r < n) {
      if (CPLX) {
        const double2 wi = reinterpret_cast<const double2*>(w)[r];
        const double ur = wi.x * inv, ui = wi.y * inv;
        const double yr = ar * inv + (shr * ur - shi * ui);
        const double yi = ai * inv + (shr * ui + shi * ur);
        reinterpret_cast<double2*>(ucol)[r] = make_double2(ur, ui);
        reinterpret_cast<double2*>(v)[r] = make_double2(yr, yi);
        d0 += ur * yr + ui * yi;
        d1 += ur * yi - ui * yr;
      } else {
        const double ui = w[r] * inv;
        const double y = ar * inv + shr * ui;
        ucol[r] = ui;
        v[r] = y;
        d0 = fma(ui, y, d0);
      }
    }
  }
  grid_sum_finalize<CPLX ? 2 : 1>(d0, d1, partial, ticket, sc.alpha_slot);
}

struct DenseOp : cmb_op {
  double* d_a = nullptr;
  ~DenseOp() override { cudaFree(d_a); }
  int apply(const double* w, double* ucol, double* v, double shr, double shi, const StepScalars& sc) override {
    // One warp per row.  Splitting a row over K = 2, 4, 8 warps fills more of the GPU at n = 2000 but measured slower
    // (16.5 us per launch with K = 1, 19.4 with 2, 21.0 with 8): the launch is bound by its fixed costs — the K-fold
    // number of CTAs in the final reduction outweighs the shorter rows.  CMPT_B200_DENSE_K overrides (measurements).
    int K = 1;
    if (const char* e = getenv("CMPT_B200_DENSE_K")) {
      const int k = atoi(e);
      if (k == 1 || k == 2 || k == 4 || k == 8) K = k;
    }
    const int rows_per_cta = 8 / K;
    int grid = int(std::min<long long>((n_local + rows_per_cta - 1) / rows_per_cta, (long long)ctx->num_sms * 8));
    if (grid < 1) grid = 1;
    LaunchScope ls(ctx, "gemv_dense");
    if (cplx)
      dense_apply_kernel<true><<<grid, 256, 0, ctx->stream>>>(d_a, n_local, w, ucol, v, shr, shi, sc, ctx->d_partial,
                                                              ctx->d_ticket + 1, K);
    else
      dense_apply_kernel<false><<<grid, 256, 0, ctx->stream>>>(d_a, n_local, w, ucol, v, shr, shi, sc, ctx->d_partial,
                                                               ctx->d_ticket + 1, K);
    CMB_CUDA(cudaGetLastError());
    return CMB_OK;
  }
};

// ======================================================================================================
// legacy host callback (lanczos.hpp:116,389,442): device -> pinned host -> user function -> device
// ======================================================================================================
__global__ void scale_step_kernel(const double* __restrict__ w, double* __restrict__ ucol, long long nd,
                                  StepScalars sc) {
  double inv;
  if (!step_prologue(sc, inv)) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) ucol[i] = w[i] * inv;
}

struct CallbackOp : cmb_op {
  cmb_matmul_fn fn = nullptr;
  void* user = nullptr;
  double* h_in = nullptr;   // pinned
  double* h_out = nullptr;  // pinned
  int* h_halt = nullptr;    // pinned
  ~CallbackOp() override {
    cudaFreeHost(h_in);
    cudaFreeHost(h_out);
    cudaFreeHost(h_halt);
  }
  int apply(const double* w, double* ucol, double* v, double shr, double shi, const StepScalars& sc) override {
    const int es = cplx ? 2 : 1;
    const long long nd = n_local * es;
    {
      LaunchScope ls(ctx, "callback_scale");
      int grid = int(std::min<long long>((nd + 255) / 256, (long long)ctx->num_sms * 8));
      scale_step_kernel<<<grid, 256, 0, ctx->stream>>>(w, ucol, nd, sc);
    }
    CMB_CUDA(cudaGetLastError());
    CMB_CUDA(cudaMemcpyAsync(h_halt, sc.halt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CMB_CUDA(cudaMemcpyAsync(h_in, ucol, sizeof(double) * nd, cudaMemcpyDeviceToHost, ctx->stream));
    CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*h_halt) return CMB_OK;
    fn(h_in, h_out, user);  // exceptions from C++ callers propagate through the header layer, not through here
    CMB_CUDA(cudaMemcpyAsync(v, h_out, sizeof(double) * nd, cudaMemcpyHostToDevice, ctx->stream));
    const long long ldp = (nd + 511) / 512 * 512;  // padded length; pads are zero
    if (shr != 0.0 || shi != 0.0) CMB_TRY(vec_axpy_shift(ctx, cplx, shr, shi, ucol, v, ldp, sc.halt));
    CMB_TRY(vec_dot(ctx, cplx, ucol, v, ldp, sc.alpha_slot, sc.halt));
    return CMB_OK;
  }
};

}  // namespace cmb

using namespace cmb;

static int op_common(cmb_op* op, cmb_ctx* ctx, cmb_dtype dtype, int64_t n_global, int64_t rb, int64_t re) {
  op->ctx = ctx;
  op->dtype = dtype;
  op->cplx = dtype == CMB_C64;
  op->n_global = n_global;
  op->row_begin = rb;
  op->n_local = re - rb;
  return CMB_OK;
}

extern "C" {

int cmb_op_csr_create(cmb_ctx* ctx, cmb_dtype dtype, int64_t n_global, int64_t row_begin, int64_t row_end,
                      const int64_t* rowptr, const int32_t* col, const void* val, cmb_op** out) {
  CMB_REQUIRE(ctx && out && rowptr, "null argument");
  *out = nullptr;
  CMB_REQUIRE(dtype == CMB_F64 || dtype == CMB_C64, "dtype must be CMB_F64 or CMB_C64");
  CMB_REQUIRE(n_global >= 0 && row_begin >= 0 && row_begin <= row_end && row_end <= n_global, "bad row range");
  CMB_REQUIRE(n_global < (int64_t(1) << 31), "column indices are 32-bit: n_global must be < 2^31");
  const int64_t n = row_end - row_begin;
  CMB_REQUIRE(rowptr[0] == 0 && rowptr[n] >= 0, "rowptr must start at 0");
  CMB_REQUIRE(rowptr[n] == 0 || (col && val), "null col/val");
  CMB_CUDA(cudaSetDevice(ctx->device));
  SellOp* op = new (std::nothrow) SellOp();
  if (!op) return CMB_ERR_NOMEM;
  op_common(op, ctx, dtype, n_global, row_begin, row_end);
  op->family = "spmv_sell";
  int rc = CMB_OK;
  // CMPT_B200_TRACE=1: wall-clock time of the build phases on stderr (each phase ends with a stream synchronisation)
  static const bool trace = getenv("CMPT_B200_TRACE") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = trace ? now() : 0.0;
  auto lap = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(ctx->stream);
    const double t1 = now();
    fprintf(stderr, "[cmpt_b200 rank %d] csr_create %-14s %8.3f ms\n", ctx->rank, what, t1 - t0);
    t0 = t1;
  };
  if (ctx->nranks > 1) {
    // row-partitioned shard: the uniform partition of SURVEY.md §8(e) is required
    if (row_begin != partition_begin(n_global, ctx->nranks, ctx->rank) ||
        row_end != partition_begin(n_global, ctx->nranks, ctx->rank + 1)) {
      set_error("rank %d must own rows [%lld,%lld) of the uniform partition", ctx->rank,
                (long long)partition_begin(n_global, ctx->nranks, ctx->rank),
                (long long)partition_begin(n_global, ctx->nranks, ctx->rank + 1));
      delete op;
      return CMB_ERR_INVALID;
    }
    // the shard goes to HBM as it is; the halo plan (distinct remote columns, remapped indices) is made there
    std::vector<int32_t> halo_cols;
    std::vector<int64_t> per_owner;
    CsrOnDevice csr;
    rc = upload_csr(op, rowptr, col, val, csr);
    lap("upload");
    if (rc == CMB_OK) rc = plan_halo_device(ctx, n_global, ctx->nranks, ctx->rank, csr.nnz, csr.col, halo_cols, per_owner);
    lap("plan_halo");
    // the halo set-up is collective: every rank enters it, whatever its local outcome so far
    op->halo = new (std::nothrow) HaloExchange();
    if (rc == CMB_OK && !op->halo) rc = CMB_ERR_NOMEM;
    if (rc != CMB_OK) {
      halo_cols.clear();
      per_owner.assign(ctx->nranks, 0);
    }
    if (op->halo) {
      const int rs = op->halo->setup(ctx, n_global, op->cplx ? 2 : 1, halo_cols, per_owner);
      if (rc == CMB_OK) rc = rs;
    }
    lap("halo_setup");
    if (rc == CMB_OK) rc = build_sell(op, csr, n + int64_t(halo_cols.size()));
    lap("build_sell");
    if (rc == CMB_OK && op->halo->p2p) rc = build_slice_order(op);
    lap("slice_order");
    rc = agree_status(ctx, rc, "cmb_op_csr_create");  // an operator exists on every rank or on none
  } else {
    CsrOnDevice csr;
    rc = upload_csr(op, rowptr, col, val, csr);
    lap("upload");
    if (rc == CMB_OK) rc = build_sell(op, csr, n_global);
    lap("build_sell");
  }
  if (rc != CMB_OK) {
    delete op;
    return rc;
  }
  const double s = op->cplx ? 16.0 : 8.0;
  // nnz*(s+idx) + (n+1)*ptr + 2*n*s with idx = ptr = 4 (+ the halo values read once)
  op->bytes = double(op->nnz) * (s + 4.0) + double(n + 1) * 4.0 + 2.0 * double(n) * s +
              (op->halo ? double(op->halo->nrecv) * s : 0.0);
  *out = op;
  return CMB_OK;
}

int cmb_op_dense_create(cmb_ctx* ctx, cmb_dtype dtype, int64_t n_global, int64_t row_begin, int64_t row_end,
                        const void* a_rows, cmb_op** out) {
  CMB_REQUIRE(ctx && out, "null argument");
  *out = nullptr;
  CMB_REQUIRE(dtype == CMB_F64 || dtype == CMB_C64, "dtype must be CMB_F64 or CMB_C64");
  CMB_REQUIRE(row_begin == 0 && row_end == n_global && n_global >= 0, "dense operators are single-rank (full rows)");
  CMB_REQUIRE(n_global == 0 || a_rows, "null matrix");
  if (ctx->nranks > 1) {
    set_error("dense operators are not row-partitioned in this build");
    return CMB_ERR_UNSUPPORTED;
  }
  CMB_CUDA(cudaSetDevice(ctx->device));
  DenseOp* op = new (std::nothrow) DenseOp();
  if (!op) return CMB_ERR_NOMEM;
  op_common(op, ctx, dtype, n_global, row_begin, row_end);
  op->family = "gemv_dense";
  const double s = op->cplx ? 16.0 : 8.0;
  const size_t bytes = size_t(double(n_global) * double(n_global) * s);
  if (cudaMalloc(&op->d_a, std::max<size_t>(bytes, 16)) != cudaSuccess) {
    cudaGetLastError();
    delete op;
    set_error("out of device memory for a %lld x %lld dense operator", (long long)n_global, (long long)n_global);
    return CMB_ERR_NOMEM;
  }
  if (bytes) {
    cudaError_t e = cudaMemcpy(op->d_a, a_rows, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      delete op;
      set_error("copying the dense operator failed: %s", cudaGetErrorString(e));
      return CMB_ERR_CUDA;
    }
  }
  op->bytes = double(n_global) * double(n_global) * s + 2.0 * double(n_global) * s;
  *out = op;
  return CMB_OK;
}

int cmb_op_callback_create(cmb_ctx* ctx, cmb_dtype dtype, int64_t n, cmb_matmul_fn fn, void* user, cmb_op** out) {
  CMB_REQUIRE(ctx && out && fn, "null argument");
  *out = nullptr;
  CMB_REQUIRE(dtype == CMB_F64 || dtype == CMB_C64, "dtype must be CMB_F64 or CMB_C64");
  CMB_REQUIRE(n >= 0, "negative height");
  if (ctx->nranks > 1) {
    set_error("host-callback operators are single-rank");
    return CMB_ERR_UNSUPPORTED;
  }
  CMB_CUDA(cudaSetDevice(ctx->device));
  CallbackOp* op = new (std::nothrow) CallbackOp();
  if (!op) return CMB_ERR_NOMEM;
  op_common(op, ctx, dtype, n, 0, n);
  op->family = "callback";
  op->fn = fn;
  op->user = user;
  const size_t bytes = std::max<size_t>(size_t(n) * (op->cplx ? 16 : 8), 16);
  if (cudaMallocHost(&op->h_in, bytes) != cudaSuccess || cudaMallocHost(&op->h_out, bytes) != cudaSuccess ||
      cudaMallocHost(&op->h_halt, sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    delete op;
    set_error("out of pinned host memory for the callback staging buffers");
    return CMB_ERR_NOMEM;
  }
  op->bytes = 2.0 * double(n) * (op->cplx ? 16.0 : 8.0);
  *out = op;
  return CMB_OK;
}

int cmb_op_destroy(cmb_op* op) {
  if (!op) return CMB_OK;
  cudaSetDevice(op->ctx->device);
  cudaStreamSynchronize(op->ctx->stream);
  delete op;
  return CMB_OK;
}

cmb_ctx* cmb_op_context(const cmb_op* op) { return op ? op->ctx : nullptr; }
int64_t cmb_op_row_begin(const cmb_op* op) { return op ? op->row_begin : 0; }
int64_t cmb_op_rows(const cmb_op* op) { return op ? op->n_local : 0; }
int64_t cmb_op_height(const cmb_op* op) { return op ? op->n_global : 0; }
int cmb_op_dtype(const cmb_op* op) { return op ? op->dtype : -1; }
double cmb_op_bytes(const cmb_op* op) { return op ? op->bytes : 0.0; }

// diagnostic (cmpt_b200_debug.h): number of halo exchanges this rank's CSR shard has completed (-1: no peer-memory halo)
int cmb_debug_op_sell_stats(cmb_op* op, long long* nnz, long long* padded, int* sorted) {
  CMB_REQUIRE(op && nnz && padded && sorted, "null argument");
  SellOp* s = dynamic_cast<SellOp*>(op);
  CMB_REQUIRE(s, "not a CSR (SELL) operator");
  *nnz = s->nnz;
  *padded = s->padded_nnz;
  *sorted = s->d_perm ? 1 : 0;
  return CMB_OK;
}

int cmb_debug_op_exchange_count(cmb_op* op, long long* count) {
  CMB_REQUIRE(op && count, "null argument");
  *count = -1;
  SellOp* s = dynamic_cast<SellOp*>(op);
  if (!s || !s->halo || !s->halo->p2p) return CMB_OK;
  CMB_CUDA(cudaSetDevice(op->ctx->device));
  CMB_CUDA(cudaStreamSynchronize(op->ctx->stream));
  unsigned long long v = 0;
  CMB_TRY(d2h_sync(op->ctx, &v, s->halo->push.xseq, sizeof(v)));
  *count = (long long)v;
  return CMB_OK;
}

int cmb_op_apply_host(cmb_op* op, const void* x, void* y) {
  CMB_REQUIRE(op && (op->n_local == 0 || (x && y)), "null argument");
  cmb_ctx* ctx = op->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  const int es = op->cplx ? 2 : 1;
  const int64_t nd = op->n_local * es;
  if (nd == 0) return CMB_OK;
  const int64_t ld = (nd + 511) / 512 * 512;
  if (ctx->dead) return check_peer_wait(ctx);
  double* buf = nullptr;
  CMB_TRY(pool_alloc(ctx, &buf, sizeof(double) * (3 * ld + 8)));  // stream-ordered pool: no cudaMalloc/cudaFree per apply
  int rc = [&]() -> int {
    double *w = buf, *u = buf + ld, *v = buf + 2 * ld, *sc = buf + 3 * ld;
    CMB_CUDA(cudaMemsetAsync(buf, 0, sizeof(double) * (3 * ld + 8), ctx->stream));
    CMB_CUDA(cudaMemcpyAsync(w, x, sizeof(double) * nd, cudaMemcpyHostToDevice, ctx->stream));
    const double one = 1.0;
    CMB_CUDA(cudaMemcpyAsync(sc, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    StepScalars s;
    s.nrm2 = sc;
    s.threshold = -1.0;
    s.halt = reinterpret_cast<int*>(sc + 1);
    s.beta_slot = sc + 2;
    s.alpha_slot = sc + 4;
    CMB_TRY(op->apply(w, u, v, 0.0, 0.0, s));
    CMB_CUDA(cudaMemcpyAsync(y, v, sizeof(double) * nd, cudaMemcpyDeviceToHost, ctx->stream));
    CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    CMB_TRY(check_peer_wait(ctx));  // a timed-out halo wait must not return stale data silently
    return CMB_OK;
  }();
  pool_free(ctx, buf);
  return rc;
}

}  // extern "C"
