// halo.cu — row-partitioned operators: halo planning (host), NVLink halo exchange (NCCL send/recv groups).
//
// The reference has no distributed code; the partition is the one §8(e) defines: rank q owns
// the contiguous global rows [q*n/P, (q+1)*n/P).  For a CSR shard the plan is
//   halo_cols   sorted distinct remote columns this rank reads (grouped by owner because owners are ranges)
//   col_local   column indices remapped to [0, n_local) for owned columns and n_local + position in halo_cols
//               for remote ones, so the SpMV kernel gathers from one of two base pointers
// and, after the owners learn what each peer needs, per apply:
//   pack (gather owned w entries into one contiguous send buffer)  ->  ncclGroup{Send,Recv per peer}  ->  SpMV.
// The values exchanged are the un-normalised w; 1/beta is a global scalar applied inside the SpMV.
#include <string.h>

#include <algorithm>
#include <thread>

#include "device_utils.cuh"
#include "halo.cuh"

namespace cmb {

int64_t partition_begin(int64_t n, int P, int q) { return (int64_t(q) * n) / P; }

static int owner_of(int64_t g, int64_t n, int P) {
  // smallest q with partition_begin(q+1) > g
  int q = int((g * P) / n);
  if (q >= P) q = P - 1;
  while (q > 0 && partition_begin(n, P, q) > g) --q;
  while (q < P - 1 && partition_begin(n, P, q + 1) <= g) ++q;
  return q;
}

int plan_halo(int64_t n, int P, int rank, int64_t nnz, const int32_t* col, int32_t* col_local,
              std::vector<int32_t>& halo_cols, std::vector<int64_t>& per_owner) {
  const int64_t r0 = partition_begin(n, P, rank), r1 = partition_begin(n, P, rank + 1);
  const int64_t nloc = r1 - r0;
  // 1. distinct remote columns: per-thread collection, then sort + unique
  const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const unsigned nt = nnz > (1 << 20) ? hw : 1;
  std::vector<std::vector<int32_t>> parts(nt);
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        const int64_t a = nnz * t / nt, b = nnz * (t + 1) / nt;
        auto& v = parts[t];
        int32_t last = -1;
        for (int64_t i = a; i < b; ++i) {
          const int32_t c = col[i];
          if ((c < r0 || c >= r1) && c != last) {
            v.push_back(c);
            last = c;
          }
        }
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
      });
    for (auto& t : th) t.join();
  }
  halo_cols.clear();
  for (auto& v : parts) halo_cols.insert(halo_cols.end(), v.begin(), v.end());
  std::sort(halo_cols.begin(), halo_cols.end());
  halo_cols.erase(std::unique(halo_cols.begin(), halo_cols.end()), halo_cols.end());
  for (int32_t c : halo_cols)
    if (c < 0 || c >= n) {
      set_error("column index %d out of range [0,%lld)", int(c), (long long)n);
      return CMB_ERR_INVALID;
    }
  per_owner.assign(P, 0);
  for (int32_t c : halo_cols) per_owner[owner_of(c, n, P)]++;
  // 2. remap
  if (col_local) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        const int64_t a = nnz * t / nt, b = nnz * (t + 1) / nt;
        for (int64_t i = a; i < b; ++i) {
          const int32_t c = col[i];
          if (c >= r0 && c < r1) {
            col_local[i] = int32_t(c - r0);
          } else {
            const auto it = std::lower_bound(halo_cols.begin(), halo_cols.end(), c);
            col_local[i] = int32_t(nloc + (it - halo_cols.begin()));
          }
        }
      });
    for (auto& t : th) t.join();
  }
  return CMB_OK;
}

// ---- the same plan on the device -----------------------------------------------------------------------------
// The host version above costs a 4-byte-per-non-zero host array (allocated, page-faulted, filled, uploaded from
// pageable memory): ~70 ms per 42 M non-zeros.  On the device the distinct remote columns are a bitmap over the global
// column range; a popcount prefix over its words gives every remote column its rank in sorted order, i.e. its position
// in the halo, without sorting.
__global__ void halo_mark_kernel(const int* __restrict__ col, long long nnz, long long r0, long long r1, long long n,
                                 unsigned* __restrict__ bitmap, int* __restrict__ bad) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    const long long c = col[i];
    if (c < 0 || c >= n) {
      *bad = 1;
    } else if (c < r0 || c >= r1) {
      const unsigned bit = 1u << (c & 31);
      unsigned* w = bitmap + (c >> 5);
      if (!(__ldg(w) & bit)) atomicOr(w, bit);  // most marks repeat: test first
    }
  }
}
// popcount of every bitmap word; block sums for the two-level exclusive scan
__global__ void halo_count_kernel(const unsigned* __restrict__ bitmap, long long nwords, unsigned* __restrict__ cnt,
                                  unsigned long long* __restrict__ block_sum) {
  __shared__ unsigned s[1024 / 32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c = i < nwords ? __popc(bitmap[i]) : 0u;
  if (i < nwords) cnt[i] = c;
  unsigned v = c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned t = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = t;
  }
}
// exclusive prefix of the block sums (one block; the number of blocks is nwords / 1024, a few hundred)
__global__ void halo_scan_blocks_kernel(unsigned long long* __restrict__ block_sum, int nblocks, unsigned long long* total) {
  if (threadIdx.x == 0) {
    unsigned long long acc = 0;
    for (int b = 0; b < nblocks; ++b) {
      const unsigned long long v = block_sum[b];
      block_sum[b] = acc;
      acc += v;
    }
    *total = acc;
  }
}
// prefix[i] = number of marked columns in words [0, i)
__global__ void halo_prefix_kernel(const unsigned* __restrict__ cnt, long long nwords, const unsigned long long* __restrict__ block_off,
                                   unsigned* __restrict__ prefix) {
  __shared__ unsigned s[1024];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  s[threadIdx.x] = i < nwords ? cnt[i] : 0u;
  __syncthreads();
  // inclusive Hillis-Steele scan over the block
  for (int o = 1; o < int(blockDim.x); o <<= 1) {
    const unsigned add = threadIdx.x >= unsigned(o) ? s[threadIdx.x - o] : 0u;
    __syncthreads();
    s[threadIdx.x] += add;
    __syncthreads();
  }
  if (i < nwords) prefix[i] = unsigned(block_off[blockIdx.x]) + s[threadIdx.x] - cnt[i];
}
__global__ void halo_remap_kernel(int* __restrict__ col, long long nnz, long long r0, long long r1, const unsigned* __restrict__ bitmap,
                                  const unsigned* __restrict__ prefix) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nloc = r1 - r0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    const long long c = col[i];
    if (c >= r0 && c < r1) {
      col[i] = int(c - r0);
    } else {
      const unsigned w = __ldg(bitmap + (c >> 5));
      col[i] = int(nloc + prefix[c >> 5] + __popc(w & ((1u << (c & 31)) - 1u)));
    }
  }
}
__global__ void halo_expand_kernel(const unsigned* __restrict__ bitmap, long long nwords, const unsigned* __restrict__ prefix,
                                   int* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nwords) return;
  unsigned w = bitmap[i];
  unsigned pos = prefix[i];
  while (w) {
    const int b = __ffs(w) - 1;
    out[pos++] = int(i * 32 + b);
    w &= w - 1;
  }
}

int plan_halo_device(cmb_ctx* ctx, int64_t n, int P, int rank, int64_t nnz, int32_t* d_col, std::vector<int32_t>& halo_cols,
                     std::vector<int64_t>& per_owner) {
  const int64_t r0 = partition_begin(n, P, rank), r1 = partition_begin(n, P, rank + 1);
  const long long nwords = (n + 31) / 32;
  const int nblocks = int((nwords + 1023) / 1024);
  unsigned *d_bitmap = nullptr, *d_cnt = nullptr, *d_prefix = nullptr;
  unsigned long long* d_bsum = nullptr;  // [nblocks] block offsets, [nblocks] total, [nblocks + 1] bad flag
  int* d_halo = nullptr;
  auto cleanup = [&]() {
    pool_free(ctx, d_bitmap);
    pool_free(ctx, d_cnt);
    pool_free(ctx, d_prefix);
    pool_free(ctx, d_bsum);
    pool_free(ctx, d_halo);
  };
  int rc = [&]() -> int {
    CMB_TRY(pool_alloc(ctx, &d_bitmap, sizeof(unsigned) * size_t(nwords)));
    CMB_TRY(pool_alloc(ctx, &d_cnt, sizeof(unsigned) * size_t(nwords)));
    CMB_TRY(pool_alloc(ctx, &d_prefix, sizeof(unsigned) * size_t(nwords)));
    CMB_TRY(pool_alloc(ctx, &d_bsum, sizeof(unsigned long long) * size_t(nblocks + 2)));
    CMB_CUDA(cudaMemsetAsync(d_bitmap, 0, sizeof(unsigned) * size_t(nwords), ctx->stream));
    CMB_CUDA(cudaMemsetAsync(d_bsum, 0, sizeof(unsigned long long) * size_t(nblocks + 2), ctx->stream));
    int* d_bad = reinterpret_cast<int*>(d_bsum + nblocks + 1);
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>((nnz + 255) / 256, int64_t(ctx->num_sms) * 16)));
    {
      LaunchScope ls(ctx, "halo_plan");
      halo_mark_kernel<<<grid, 256, 0, ctx->stream>>>(d_col, nnz, r0, r1, n, d_bitmap, d_bad);
      halo_count_kernel<<<nblocks, 1024, 0, ctx->stream>>>(d_bitmap, nwords, d_cnt, d_bsum);
      halo_scan_blocks_kernel<<<1, 32, 0, ctx->stream>>>(d_bsum, nblocks, d_bsum + nblocks);
      halo_prefix_kernel<<<nblocks, 1024, 0, ctx->stream>>>(d_cnt, nwords, d_bsum, d_prefix);
    }
    CMB_CUDA(cudaGetLastError());
    unsigned long long tail[2] = {0, 0};
    CMB_TRY(d2h_sync(ctx, tail, d_bsum + nblocks, sizeof(tail)));
    int bad = 0;
    memcpy(&bad, &tail[1], sizeof(int));
    if (bad) {
      set_error("column index out of range [0,%lld)", (long long)n);
      return CMB_ERR_INVALID;
    }
    const int64_t nhalo = int64_t(tail[0]);
    halo_cols.resize(size_t(nhalo));
    CMB_TRY(pool_alloc(ctx, &d_halo, sizeof(int) * size_t(std::max<int64_t>(nhalo, 1))));
    {
      LaunchScope ls(ctx, "halo_plan");
      halo_remap_kernel<<<grid, 256, 0, ctx->stream>>>(d_col, nnz, r0, r1, d_bitmap, d_prefix);
      halo_expand_kernel<<<nblocks, 1024, 0, ctx->stream>>>(d_bitmap, nwords, d_prefix, d_halo);
    }
    CMB_CUDA(cudaGetLastError());
    if (nhalo) CMB_TRY(d2h_sync(ctx, halo_cols.data(), d_halo, sizeof(int) * size_t(nhalo)));
    else CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    // owners are ranges of the sorted list
    per_owner.assign(P, 0);
    for (int q = 0; q < P; ++q) {
      const auto lo = std::lower_bound(halo_cols.begin(), halo_cols.end(), int32_t(partition_begin(n, P, q)));
      const auto hi = std::lower_bound(halo_cols.begin(), halo_cols.end(), int32_t(std::min<int64_t>(partition_begin(n, P, q + 1), INT32_MAX)));
      per_owner[q] = int64_t(hi - lo);
    }
    return CMB_OK;
  }();
  cleanup();
  return rc;
}

// ---- NCCL helpers --------------------------------------------------------------------------------------
static int nccl_check(cmb_ctx* ctx, int r, const char* what) {
  if (r != 0) {
    set_error("%s failed: %s", what, ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "?");
    return CMB_ERR_NCCL;
  }
  return CMB_OK;
}

template <int ES>
__global__ void pack_kernel(const double* __restrict__ w, const int* __restrict__ idx, long long count,
                            double* __restrict__ out, const int* __restrict__ halt) {
  if (*halt) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const long long s = idx[i];
#pragma unroll
    for (int e = 0; e < ES; ++e) out[i * ES + e] = w[s * ES + e];
  }
}

// Collective on multi-rank contexts: a peer may still have this rank's receive buffer mapped (or be pushing into it),
// so every rank first unmaps, then all ranks meet, and only then the exported buffer is freed.
HaloExchange::~HaloExchange() {
  if (p2p_ctx) {
    cudaStreamSynchronize(p2p_ctx->stream);
    ipc_unshare(p2p_ctx, p2p_mapped);
    rank_barrier(p2p_ctx);
    dfree(p2p_ctx, push.xseq);
    dfree(p2p_ctx, push.ticket);
    dfree(p2p_ctx, p2p_base);
  }
  dfree(own_ctx, d_send_idx);
  dfree(own_ctx, d_sendbuf);
  dfree(own_ctx, d_halo);
}

// Collective.  cnt is the P x P matrix of halo counts (cnt[q*P + r] = entries rank q reads from rank r).
int HaloExchange::setup_p2p(cmb_ctx* ctx, const std::vector<double>& cnt) {
  if (!ctx->mail_ok || getenv("CMPT_B200_NO_P2P_HALO")) return CMB_OK;  // mail_ok: IPC between all ranks works
  constexpr size_t kFlagBytes = 256;
  static_assert(kMaxPeers * sizeof(unsigned long long) <= kFlagBytes, "flag block too small");
  const size_t bytes = kFlagBytes + 2 * sizeof(double) * size_t(es) * size_t(std::max<int64_t>(nrecv, 1));
  void* base = nullptr;
  if (cudaMalloc(&base, bytes) != cudaSuccess) {
    cudaGetLastError();
    base = nullptr;
  } else {
    cudaMemsetAsync(base, 0, bytes, ctx->stream);
  }
  unsigned long long* xseq = nullptr;
  unsigned* ticket = nullptr;
  bool ok = base && cudaMalloc(&xseq, sizeof(unsigned long long)) == cudaSuccess &&
            cudaMalloc(&ticket, sizeof(unsigned)) == cudaSuccess;
  if (ok) {
    cudaMemsetAsync(xseq, 0, sizeof(unsigned long long), ctx->stream);
    cudaMemsetAsync(ticket, 0, sizeof(unsigned), ctx->stream);
  }
  cudaStreamSynchronize(ctx->stream);
  void* mapped[kMaxPeers];
  if (!ipc_share(ctx, ok ? base : nullptr, mapped)) {
    dfree(ctx, base);
    dfree(ctx, xseq);
    dfree(ctx, ticket);
    cudaGetLastError();
    return CMB_OK;  // NCCL path stays
  }
  p2p = true;
  p2p_ctx = ctx;
  p2p_base = base;
  push.P = P;
  push.rank = rank;
  push.xseq = xseq;
  push.ticket = ticket;
  push.idx = d_send_idx;
  push.npush = 1;  // set per launch (fused_push)
  for (int q = 0; q <= P; ++q) push.send_off[q] = send_off[q];
  for (int q = 0; q < P; ++q) {
    p2p_mapped[q] = mapped[q];
    int64_t off = 0, tot = 0;
    for (int r = 0; r < P; ++r) {
      if (r < rank) off += int64_t(cnt[size_t(q) * P + r]);
      tot += int64_t(cnt[size_t(q) * P + r]);
    }
    push.dst[q] = reinterpret_cast<double*>(static_cast<char*>(mapped[q]) + kFlagBytes) + off * es;
    push.stride[q] = std::max<int64_t>(tot, 1) * es;
    push.flag[q] = static_cast<unsigned long long*>(mapped[q]) + rank;
  }
  pull.base = reinterpret_cast<const double*>(static_cast<char*>(base) + kFlagBytes);
  pull.stride = std::max<int64_t>(nrecv, 1) * es;
  pull.flag = static_cast<const unsigned long long*>(base);
  pull.xseq = xseq;
  pull.P = P;
  pull.rank = rank;
  pull.error = ctx->d_mail_error;
  pull.timeout = ctx->spin_timeout;
  return CMB_OK;
}

int HaloExchange::setup(cmb_ctx* ctx, int64_t n, int es_, const std::vector<int32_t>& halo_cols,
                        const std::vector<int64_t>& per_owner) {
  P = ctx->nranks;
  rank = ctx->rank;
  own_ctx = ctx;
  es = es_;
  const int64_t r0 = partition_begin(n, P, rank);
  recv_off.assign(P + 1, 0);
  for (int q = 0; q < P; ++q) recv_off[q + 1] = recv_off[q] + per_owner[q];
  nrecv = recv_off[P];
  // 1. everybody learns the P x P matrix of counts: allreduce of a matrix in which each rank fills its row
  double* d_cnt = nullptr;
  CMB_CUDA(cudaMalloc(&d_cnt, sizeof(double) * P * P));
  std::vector<double> cnt(size_t(P) * P, 0.0);
  for (int q = 0; q < P; ++q) cnt[size_t(rank) * P + q] = double(per_owner[q]);
  CMB_CUDA(cudaMemcpyAsync(d_cnt, cnt.data(), sizeof(double) * P * P, cudaMemcpyHostToDevice, ctx->stream));
  int rc = allreduce_sum_f64(ctx, d_cnt, size_t(P) * P);
  if (rc == CMB_OK) {
    cudaError_t e = cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(double) * P * P, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("halo setup: %s", cudaGetErrorString(e));
      rc = CMB_ERR_CUDA;
    }
  }
  dfree(ctx, d_cnt);
  CMB_TRY(rc);
  send_off.assign(P + 1, 0);
  for (int q = 0; q < P; ++q) send_off[q + 1] = send_off[q] + int64_t(cnt[size_t(q) * P + rank]);  // what q needs from me
  nsend = send_off[P];
  // 2. exchange the index lists (global column numbers), then localise the ones I have to serve
  int32_t* d_need = nullptr;
  CMB_CUDA(cudaMalloc(&d_need, sizeof(int32_t) * std::max<int64_t>(nrecv, 1)));
  CMB_CUDA(cudaMalloc(&d_send_idx, sizeof(int32_t) * std::max<int64_t>(nsend, 1)));
  CMB_CUDA(cudaMemcpyAsync(d_need, halo_cols.data(), sizeof(int32_t) * nrecv, cudaMemcpyHostToDevice, ctx->stream));
  // what I need from q goes to q; what q needs from me arrives in d_send_idx
  rc = alltoallv_i32(ctx, d_need, recv_off.data(), d_send_idx, send_off.data());
  std::vector<int32_t> sidx(std::max<int64_t>(nsend, 1));
  if (rc == CMB_OK) {
    cudaError_t e = cudaMemcpyAsync(sidx.data(), d_send_idx, sizeof(int32_t) * nsend, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("halo setup: %s", cudaGetErrorString(e));
      rc = CMB_ERR_CUDA;
    }
  }
  dfree(ctx, d_need);
  CMB_TRY(rc);
  const int64_t nloc = partition_begin(n, P, rank + 1) - r0;
  // rank-local checks and copies: their outcome is agreed on before the next collective step (the peer-memory mapping)
  rc = [&]() -> int {
    for (int64_t i = 0; i < nsend; ++i) {
      const int64_t l = int64_t(sidx[i]) - r0;
      if (l < 0 || l >= nloc) {
        set_error("halo setup: peer asked for row %d which rank %d does not own", int(sidx[i]), rank);
        return CMB_ERR_INVALID;
      }
      sidx[i] = int32_t(l);
    }
    CMB_CUDA(cudaMemcpyAsync(d_send_idx, sidx.data(), sizeof(int32_t) * nsend, cudaMemcpyHostToDevice, ctx->stream));
    CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return CMB_OK;
  }();
  CMB_TRY(agree_status(ctx, rc, "halo setup"));
  CMB_TRY(setup_p2p(ctx, cnt));
  if (!p2p) {
    CMB_CUDA(cudaMalloc(&d_sendbuf, sizeof(double) * es * std::max<int64_t>(nsend, 1)));
    CMB_CUDA(cudaMalloc(&d_halo, sizeof(double) * es * std::max<int64_t>(nrecv, 1)));
  }
  return CMB_OK;
}

int HaloExchange::exchange(cmb_ctx* ctx, const double* w, const int* halt) {
  if (p2p) return CMB_OK;  // the push is part of the consuming kernel (fused_push)
  if (ctx->vgroup) {
    set_error("virtual ranks exchange halos through peer memory only (CMPT_B200_NO_P2P_HALO is not supported there)");
    return CMB_ERR_UNSUPPORTED;
  }
  if (nsend > 0) {
    LaunchScope ls(ctx, "halo_pack");
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>((nsend + 255) / 256, int64_t(ctx->num_sms) * 8)));
    if (es == 2)
      pack_kernel<2><<<grid, 256, 0, ctx->stream>>>(w, d_send_idx, nsend, d_sendbuf, halt);
    else
      pack_kernel<1><<<grid, 256, 0, ctx->stream>>>(w, d_send_idx, nsend, d_sendbuf, halt);
    CMB_CUDA(cudaGetLastError());
  }
  LaunchScope ls(ctx, "nccl_halo");
  int rc = nccl_check(ctx, ctx->nccl->GroupStart(), "ncclGroupStart");
  for (int q = 0; q < P && rc == CMB_OK; ++q) {
    if (q == rank) continue;
    if (send_off[q + 1] > send_off[q])
      rc = nccl_check(ctx, ctx->nccl->Send(d_sendbuf + send_off[q] * es, size_t(send_off[q + 1] - send_off[q]) * es,
                                           kNcclFloat64, q, ctx->nccl_comm, ctx->stream), "ncclSend");
    if (rc == CMB_OK && recv_off[q + 1] > recv_off[q])
      rc = nccl_check(ctx, ctx->nccl->Recv(d_halo + recv_off[q] * es, size_t(recv_off[q + 1] - recv_off[q]) * es,
                                           kNcclFloat64, q, ctx->nccl_comm, ctx->stream), "ncclRecv");
  }
  if (rc == CMB_OK) rc = nccl_check(ctx, ctx->nccl->GroupEnd(), "ncclGroupEnd");
  else ctx->nccl->GroupEnd();
  return rc;
}

}  // namespace cmb

using namespace cmb;

extern "C" {

int64_t cmb_partition_begin(int64_t n_global, int nranks, int rank) {
  if (nranks <= 0 || rank < 0 || rank > nranks || n_global < 0) return -1;
  return partition_begin(n_global, nranks, rank);
}

int cmb_plan_halo(int64_t n_global, int nranks, int rank, int64_t nnz, const int32_t* col, int32_t* col_local,
                  int64_t* halo_count, int64_t* per_owner_counts, int32_t* halo_cols_out, int64_t halo_capacity) {
  CMB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks && n_global >= 0 && nnz >= 0, "bad partition");
  CMB_REQUIRE(nnz == 0 || col, "null column array");
  std::vector<int32_t> halo;
  std::vector<int64_t> per;
  CMB_TRY(plan_halo(n_global, nranks, rank, nnz, col, col_local, halo, per));
  if (halo_count) *halo_count = int64_t(halo.size());
  if (per_owner_counts) std::copy(per.begin(), per.end(), per_owner_counts);
  if (halo_cols_out) {
    CMB_REQUIRE(halo_capacity >= int64_t(halo.size()), "halo_cols capacity too small");
    std::copy(halo.begin(), halo.end(), halo_cols_out);
  }
  return CMB_OK;
}

}  // extern "C"
