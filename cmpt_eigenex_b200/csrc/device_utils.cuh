// device_utils.cuh — device helpers shared by the operator-apply and vector kernels.
#pragma once
#include <stdio.h>
#include "kernels.cuh"

namespace cmb {

// Every operator apply starts with the same prologue (LanczosBase::updateLanczosSteps, lanczos.hpp:429-439):
// beta = ||w|| ; if beta <= threshold the step is dropped ; else u = w / beta.  The decision is taken on the
// device so that a chain of steps can be enqueued without a host round trip; `halt` is sticky.
// ---- peer-memory mailboxes (see common.cuh) ---------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Wait (whole warp) until every rank has published sequence number m.seq in this rank's mailbox.  The spin is
// bounded (m.timeout SM clocks, ~30 s by default): a missing peer raises m.error and the sticky halt flag, so that no
// later launch of the chain consumes stale data, instead of hanging the GPU.
__device__ __forceinline__ void mail_wait(const MailPull& m) {
  const int lane = threadIdx.x & 31;
  if (lane < m.P) {
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(m.flag + lane) < m.seq) {
      if (clock64() - t0 > m.timeout) {
        if (*m.error == 0)
          printf("libcmpt_b200: mailbox wait timed out: sender %d has published %llu, waiting for %llu (block %d)\n", lane,
                 ld_acquire_sys_u64(m.flag + lane), m.seq, int(blockIdx.x));
        *m.error = 1;
        if (m.halt) *m.halt = 1;
        break;
      }
    }
  }
  __syncwarp();
}
// Sum of the P partials of entry idx, always in rank order: every rank gets identical bits.
__device__ __forceinline__ double mail_sum(const MailPull& m, int idx) {
  double s = 0.0;
  for (int q = 0; q < m.P; ++q) s += ld_relaxed_sys_f64(m.data + q * kMailStride + idx);
  return s;
}
// Push value `v` as entry idx of this rank into every rank's mailbox (NVLink stores).
__device__ __forceinline__ void mail_push_value(const MailPush& m, int idx, double v) {
  for (int q = 0; q < m.P; ++q) st_relaxed_sys_f64(m.data[q] + m.rank * kMailStride + idx, v);
}
// Publish: called by the first m.P threads of the CTA (one peer each, so that the P release stores and their NVLink
// round trips overlap) after all pushing threads fenced (__threadfence_system) and synchronised.
__device__ __forceinline__ void mail_publish(const MailPush& m, int q) {
  if (q < m.P) st_release_sys_u64(m.flag[q] + m.rank, m.seq);
}

// ---- halo exchange through peer memory, fused into the operator kernel (halo.cu sets the arguments up) -------
// Producer part, executed by the CTAs with blockIdx.x < hp.npush at the start of the kernel: the values the peers
// need go straight into their receive buffers (NVLink stores); the pusher that finishes last raises this rank's
// flag at every peer (one peer per thread).  Must be called by all threads of those CTAs.
template <int ES>
__device__ __forceinline__ void halo_push_part(const HaloPush& hp, const double* __restrict__ w, unsigned long long seq) {
  const long long stride = (long long)hp.npush * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int q = 0; q < hp.P; ++q) {
    const long long a = hp.send_off[q], b = hp.send_off[q + 1];
    if (b <= a) continue;
    double* dst = hp.dst[q] + (seq & 1ull) * hp.stride[q];
    for (long long i = a + t; i < b; i += stride) {
      const long long s = hp.idx[i];
#pragma unroll
      for (int e = 0; e < ES; ++e) dst[(i - a) * ES + e] = w[s * ES + e];
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_push_last;
#ifdef CMB_HALO_TRACE
  unsigned trace_prev = 0;
  if (threadIdx.x == 0) {
    trace_prev = atomicAdd(hp.ticket, 1u);
    s_push_last = (trace_prev == unsigned(hp.npush) - 1u);
    printf("TRACE push rank %d seq %llu block %d ticket_before %u npush %d all %d last %d flag0 %p flag1 %p\n", hp.rank, seq,
           int(blockIdx.x), trace_prev, hp.npush, hp.all_push, s_push_last, (void*)hp.flag[0], (void*)hp.flag[1]);
  }
#else
  if (threadIdx.x == 0) s_push_last = (atomicAdd(hp.ticket, 1u) == unsigned(hp.npush) - 1u);
#endif
  __syncthreads();
  if (s_push_last) {
    if (int(threadIdx.x) < hp.P && int(threadIdx.x) != hp.rank) st_release_sys_u64(hp.flag[threadIdx.x], seq);
    if (threadIdx.x == 0) *hp.ticket = 0u;
  }
}
// Consumer part (all threads of the CTA): wait until every peer has published exchange `seq`.  One warp polls (with
// a short back-off, so that thousands of waiting warps do not compete with the incoming NVLink writes for L2), the
// others park on the CTA barrier.  Bounded like mail_wait.
__device__ __forceinline__ void halo_wait_cta(const HaloPull& h, unsigned long long seq) {
#ifdef CMB_HALO_TRACE
  if (threadIdx.x == 0)
    printf("TRACE wait rank %d seq %llu block %d flags@%p = %llu %llu\n", h.rank, seq, int(blockIdx.x), (void*)h.flag,
           ld_acquire_sys_u64(h.flag), ld_acquire_sys_u64(h.flag + 1));
#endif
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (lane < h.P && lane != h.rank) {
      const long long t0 = clock64();
      while (ld_acquire_sys_u64(h.flag + lane) < seq) {
        __nanosleep(100);
        if (clock64() - t0 > h.timeout) {
          if (*h.error == 0)
            printf("libcmpt_b200: halo wait timed out on rank %d: sender %d has published exchange %llu, waiting for %llu "
                   "(block %d)\n", h.rank, lane, ld_acquire_sys_u64(h.flag + lane), seq, int(blockIdx.x));
          *h.error = 1;
          if (h.halt) *h.halt = 1;
          break;
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ bool step_prologue(const StepScalars& sc, double& inv) {
  if (*sc.halt) return false;
  double nrm2;
  if (sc.nrm2_pull.P > 1) {
    mail_wait(sc.nrm2_pull);
    nrm2 = mail_sum(sc.nrm2_pull, 0);
    if (blockIdx.x == 0 && threadIdx.x == 0 && sc.nrm2_pull.writeback) *sc.nrm2_pull.writeback = nrm2;
  } else {
    nrm2 = *sc.nrm2;
  }
  const double beta = sqrt(nrm2);
  const bool first = (blockIdx.x == 0 && threadIdx.x == 0);
  if (first && sc.beta_slot) *sc.beta_slot = beta;
  if (beta <= sc.threshold) {
    if (first) *sc.halt = 1;
    return false;
  }
  inv = 1.0 / beta;
  return true;
}

// Block-level sum of `nval` (1 or 2) per-thread doubles, per-CTA partial, last-CTA deterministic final sum
// written to out[0..nval).  Must be called by every thread of every CTA of the grid exactly once.
template <int NVAL>
__device__ __forceinline__ bool grid_sum_finalize(double a0, double a1, double* partial, unsigned* ticket,
                                                  double* out) {
  __shared__ double s_red[2][32];
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  a0 = warp_sum(a0);
  if (NVAL > 1) a1 = warp_sum(a1);
  if (lane == 0) {
    s_red[0][warp] = a0;
    if (NVAL > 1) s_red[1][warp] = a1;
  }
  __syncthreads();
  if (warp == 0) {
    double b0 = lane < nw ? s_red[0][lane] : 0.0;
    double b1 = (NVAL > 1 && lane < nw) ? s_red[1][lane] : 0.0;
    b0 = warp_sum(b0);
    if (NVAL > 1) b1 = warp_sum(b1);
    if (lane == 0) {
      partial[size_t(blockIdx.x) * 2] = b0;
      if (NVAL > 1) partial[size_t(blockIdx.x) * 2 + 1] = b1;
      __threadfence();
      const unsigned prev = atomicAdd(ticket, 1u);
      s_last = (prev == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    double c0 = 0.0, c1 = 0.0;
    for (int b = lane; b < int(gridDim.x); b += 32) {
      c0 += __ldcg(&partial[size_t(b) * 2]);
      if (NVAL > 1) c1 += __ldcg(&partial[size_t(b) * 2 + 1]);
    }
    c0 = warp_sum(c0);
    if (NVAL > 1) c1 = warp_sum(c1);
    if (lane == 0) {
      out[0] = c0;
      if (NVAL > 1) out[1] = c1;
      *ticket = 0u;
    }
  }
  return s_last != 0;  // true in the CTA that finished last (every other CTA has left the kernel body by then)
}

}  // namespace cmb
