// cgs.cu — fused tall-skinny Gram-Schmidt passes over the device-resident Krylov basis.
//
// Replaces the dot/axpy pairs of LanczosBase::orthogonalize (lanczos.hpp:143-146, called c times per
// step at :416-418) and ArnoldiBase's MGS loop (arnoldi.hpp:380-383) by classical Gram-Schmidt with
// reorthogonalisation (CGS2) in three passes, each streaming the basis once:
//   DOT          h  = V^H x
//   UPDATE_DOT   y  = x - V hin ;  h = V^H y      (both from the same shared-memory tile)
//   UPDATE_NORM  y  = x - V hin ;  nrm2 = ||y||^2
// HBM-bound (0.25 flop/B): no tensor cores.  One persistent CTA per SM; one producer thread feeds a
// ring of shared-memory stages with TMA (2D tiled loads of a [T rows x nc cols] box of V, 1D bulk
// copy of the x tile); 8 consumer warps each own a column group (cg columns held in registers for
// both the update and the dot) and a row group; per-column sums live in registers across all tiles of
// the CTA and are reduced once at the end: warp shuffle -> shared memory -> per-CTA partial -> the last
// CTA to finish sums the partials in CTA order (deterministic, no floating-point atomics).
//
// The number of columns per warp `cg` is a run-time value (<= the compile-time register budget CG), so the
// TMA box holds exactly WC*cg >= c columns: shared-memory traffic (the limiter of the UPDATE passes,
// profiles/r1_prof_cgs_raw.txt) is spent on real columns only.
//
// All vectors are double arrays of padded length ld (multiple of 512, pads are zero), a complex vector
// being ld/2 interleaved (re,im) pairs; a lane always holds one double2 per column, i.e. one complex
// element or two real rows.
#include <algorithm>

#include "common.cuh"
#include "device_utils.cuh"
#include "kernels.cuh"

namespace cmb {

constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;

template <int WC>
struct CgsCfg {
  static constexpr int WR = kConsumerWarps / WC;
  static constexpr int T = 64 * WR;               // rows (doubles) per tile
  static constexpr int BOXR = T < 256 ? T : 256;  // TMA box rows
  static constexpr int NBOX = T / BOXR;
  static constexpr int X_BYTES = T * 8;
  static constexpr int PW_BYTES = (WC > 1) ? 2 * WC * T * 8 : 0;
};

template <int CG, int WC, bool CPLX, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
cgs_kernel(const __grid_constant__ CUtensorMap tmV, const double* __restrict__ x, double* __restrict__ y,
           const double* __restrict__ hin, double* __restrict__ hout, double* __restrict__ partial,
           unsigned* __restrict__ ticket, const int* __restrict__ halt, int ncols, int cg, int ntiles, int stages,
           MailPull pull, MailPush push, int norm_trick, int* retry, int retry_tag, double norm_guard,
           const __grid_constant__ SlabPush slab, int v_stable) {
  using Cfg = CgsCfg<WC>;
  constexpr int T = Cfg::T, WR = Cfg::WR;
  constexpr int ES = CPLX ? 2 : 1;  // doubles per coefficient
  constexpr int NCMAX = CG * WC;
  const int nc = cg * WC;                  // columns in the TMA box (>= ncols)
  const int stage_doubles = nc * T + T;    // V tile + x tile

  extern __shared__ __align__(1024) unsigned char smem[];
  double* stage_base = reinterpret_cast<double*>(smem);
  unsigned char* tail = smem + size_t(stages) * stage_doubles * 8;
  double* pw = reinterpret_cast<double*>(tail);
  double* red = reinterpret_cast<double*>(tail + Cfg::PW_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(tail + Cfg::PW_BYTES + WR * NCMAX * 2 * 8);
  uint64_t* empty = full + stages;
  __shared__ int s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ===== producer: one thread issues every TMA of this CTA =====
    if (lane == 0) {
      prefetch_tmap(&tmV);
      const uint32_t bytes = uint32_t(nc) * T * 8 + (x ? Cfg::X_BYTES : 0);
      auto load_v = [&](int tile, int s) {
        double* vs = stage_base + size_t(s) * stage_doubles;
#pragma unroll
        for (int b = 0; b < Cfg::NBOX; ++b)
          tma_load_2d(vs + size_t(b) * nc * Cfg::BOXR, &tmV, tile * T + b * Cfg::BOXR, 0, &full[s]);
      };
      auto load_x = [&](int tile, int s) {
        double* vs = stage_base + size_t(s) * stage_doubles;
        if (x) bulk_load_1d(vs + nc * T, x + size_t(tile) * T, Cfg::X_BYTES, &full[s]);
      };
      // The basis tiles of the first ring round do not depend on the kernel before this one (v_stable): they are
      // fetched while it drains.  Everything else waits for it.
      int npre = 0;
      if (v_stable)
        for (int tile = blockIdx.x; tile < ntiles && npre < stages; tile += gridDim.x, ++npre) {
          mbar_arrive_expect_tx(&full[npre], bytes);
          load_v(tile, npre);
        }
      grid_dependency_wait();
      grid_launch_dependents();
      const bool halted = *halt != 0;
      int tile = blockIdx.x;
      for (int i = 0; i < npre; ++i, tile += gridDim.x) load_x(tile, i);  // (stale on a halted chain, never read)
      if (halted) {
        for (int i = 0; i < npre; ++i) mbar_wait(&full[i], 0);  // no copy may be in flight when the CTA leaves
        return;
      }
      int s = npre == stages ? 0 : npre;
      uint32_t ph = npre == stages ? 1u : 0u;  // stage index / ring phase kept incrementally
      for (; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], bytes);
        load_v(tile, s);
        load_x(tile, s);
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    } else {
      grid_dependency_wait();
      grid_launch_dependents();
    }
    return;
  }
  grid_dependency_wait();
  grid_launch_dependents();
  if (*halt) return;

  // ===== consumers =====
  const int gc = warp % WC;  // column group
  const int gr = warp / WC;  // row group (64 doubles each)
  const int row = gr * 64 + 2 * lane;  // first of this lane's two doubles in the tile
  const int vofs = (row / Cfg::BOXR) * nc * Cfg::BOXR + (row % Cfg::BOXR) + gc * cg * Cfg::BOXR;

  double hr[CG], hi[CPLX ? CG : 1];
  if (MODE >= 1) {
    const bool mailed = pull.P > 1;  // coefficients = sum over ranks of the partials pushed into the mailbox
    if (mailed) mail_wait(pull);
#pragma unroll
    for (int j = 0; j < CG; ++j) {
      const int col = gc * cg + j;
      const bool ok = (j < cg) && (col < ncols);
      if (mailed) {
        hr[j] = ok ? mail_sum(pull, col * ES) : 0.0;
        if (CPLX) hi[j] = ok ? mail_sum(pull, col * ES + 1) : 0.0;
        if (ok && blockIdx.x == 0 && gr == 0 && lane == 0 && pull.writeback) {
          pull.writeback[col * ES] = hr[j];
          if (CPLX) pull.writeback[col * ES + 1] = hi[j];
        }
      } else {
        hr[j] = ok ? hin[col * ES] : 0.0;
        if (CPLX) hi[j] = ok ? hin[col * ES + 1] : 0.0;
      }
    }
  }
  double ar[CG], ai[CPLX ? CG : 1];
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    ar[j] = 0.0;
    if (CPLX) ai[j] = 0.0;
  }
  double nrm = 0.0;

  // Columns j >= cg of this warp's share do not exist: their v stays zero (set once here, the loads below are
  // predicated) and their coefficients are zero, so the arithmetic of the tile loop runs unconditionally over all CG
  // slots — no per-column select or branch in the hot loop.
  double2 v[CG];
#pragma unroll
  for (int j = 0; j < CG; ++j) v[j] = make_double2(0.0, 0.0);

  int it = 0, s = 0;
  uint32_t ph = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    mbar_wait(&full[s], ph);
    const double* vs = stage_base + size_t(s) * stage_doubles;
    double2 xv = make_double2(0.0, 0.0);
    if (x) xv = *reinterpret_cast<const double2*>(vs + nc * T + row);
#pragma unroll
    for (int j = 0; j < CG; ++j)
      if (j < cg) v[j] = *reinterpret_cast<const double2*>(vs + vofs + j * Cfg::BOXR);

    double2 yv = xv;
    if (MODE >= 1) {
      // two independent accumulation chains per component (even / odd columns) halve the dependent-FMA latency
      double2 p = make_double2(0.0, 0.0), p2 = make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < CG; ++j) {
        double2& q = (j & 1) ? p2 : p;
        if (CPLX) {
          q.x = fma(v[j].x, hr[j], q.x);
          q.x = fma(-v[j].y, hi[j], q.x);
          q.y = fma(v[j].x, hi[j], q.y);
          q.y = fma(v[j].y, hr[j], q.y);
        } else {
          q.x = fma(v[j].x, hr[j], q.x);
          q.y = fma(v[j].y, hr[j], q.y);
        }
      }
      p.x += p2.x;
      p.y += p2.y;
      if (WC > 1) {
        double* buf = pw + size_t(it & 1) * WC * T;
        *reinterpret_cast<double2*>(buf + gc * T + row) = p;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + gr), "r"(WC * 32) : "memory");
        double2 q[WC];
#pragma unroll
        for (int g = 0; g < WC; ++g) q[g] = *reinterpret_cast<const double2*>(buf + g * T + row);
        // fixed-shape pairwise tree (same order in every warp: all warps of a row group get identical y)
#pragma unroll
        for (int w = 1; w < WC; w <<= 1)
#pragma unroll
          for (int g = 0; g + w < WC; g += 2 * w) {
            q[g].x += q[g + w].x;
            q[g].y += q[g + w].y;
          }
        p = q[0];
      }
      yv.x = xv.x - p.x;
      yv.y = xv.y - p.y;
      if (gc == 0) {
        *reinterpret_cast<double2*>(y + size_t(tile) * T + row) = yv;
        if (MODE == 2 && slab.n > 0) {  // the same two doubles go to the partner ranks that need them (NVLink stores)
          const long long e = (long long)tile * T + row;
          if (e < slab.nd) {
            for (int d = 0; d < slab.n; ++d) {
              if (slab.kind[d] == 0) {
                if (e >= slab.lo[d] && e < slab.hi[d]) *reinterpret_cast<double2*>(slab.dst[d] + (e - slab.lo[d])) = yv;
              } else if (CPLX) {
                const long long idx = e >> 1;  // this lane holds one complex element
                if ((idx & 1) == slab.lo[d]) *reinterpret_cast<double2*>(slab.dst[d] + (idx >> 1) * 2) = yv;
              } else {
                slab.dst[d][e >> 1] = slab.lo[d] ? yv.y : yv.x;  // one of the lane's two real elements
              }
            }
          }
        }
      }
    }
    if (MODE <= 1) {
#pragma unroll
      for (int j = 0; j < CG; ++j) {
        if (CPLX) {  // conj(v) * y
          ar[j] = fma(v[j].x, yv.x, ar[j]);
          ar[j] = fma(v[j].y, yv.y, ar[j]);
          ai[j] = fma(v[j].x, yv.y, ai[j]);
          ai[j] = fma(-v[j].y, yv.x, ai[j]);
        } else {
          ar[j] = fma(v[j].x, yv.x, ar[j]);
          ar[j] = fma(v[j].y, yv.y, ar[j]);
        }
      }
      if (MODE == 1 && norm_trick && gc == 0) {
        nrm = fma(yv.x, yv.x, nrm);
        nrm = fma(yv.y, yv.y, nrm);
      }
    } else if (gc == 0 && !norm_trick) {
      nrm = fma(yv.x, yv.x, nrm);
      nrm = fma(yv.y, yv.y, nrm);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == stages) {
      s = 0;
      ph ^= 1u;
    }
  }

  // ===== slab push: the CTA that finishes last publishes the exchange in the partners' flags =====
  if (MODE == 2 && slab.n > 0) {
    __threadfence_system();
    asm volatile("bar.sync 9, %0;" ::"r"(kConsumerWarps * 32) : "memory");
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(slab.ticket, 1u);
      if (prev == gridDim.x - 1) {
        *slab.ticket = 0u;
        __threadfence_system();
        for (int d = 0; d < slab.n; ++d) st_release_sys_u64(slab.flag[d], slab.seq);
      }
    }
  }
  // ===== UPDATE_NORM without a reduction of its own (norm_trick): beta^2 from the reduced coefficients =====
  __shared__ double s_fin[2 * kConsumerWarps * 32 + 8];
  if (MODE == 2 && norm_trick) {
    if (blockIdx.x != 0) return;
    const int nh = ncols * ES;
    for (int t = threadIdx.x; t <= nh; t += kConsumerWarps * 32)
      s_fin[t] = (pull.P > 1) ? mail_sum(pull, t) : hin[t];
    asm volatile("bar.sync 9, %0;" ::"r"(kConsumerWarps * 32) : "memory");
    if (threadIdx.x == 0) {
      double acc = 0.0;
      for (int t = 0; t < nh; ++t) acc = fma(s_fin[t], s_fin[t], acc);
      const double b2 = s_fin[nh] - acc;
      hout[0] = b2 > 0.0 ? b2 : 0.0;
      // cancellation guard (same bits on every rank, so all ranks halt together): the host reduces the norm
      // explicitly and resumes the chain.  A vanishing ||y||^2 is a genuine breakdown, not a cancellation.
      if (retry && s_fin[nh] > 0.0 && !(b2 > norm_guard * s_fin[nh])) {
        retry[0] = 1;
        retry[1] = retry_tag;
        *const_cast<int*>(halt) = 1;
      }
    }
    return;
  }

  // ===== CTA-level reduction =====
  const int nred = (MODE <= 1) ? nc * ES : 1;  // values per CTA
  __shared__ double s_nrm[kConsumerWarps];
  if (MODE <= 1) {
#pragma unroll
    for (int j = 0; j < CG; ++j) {
      if (j < cg) {
        const double sr = warp_sum(ar[j]);
        double si = 0.0;
        if (CPLX) si = warp_sum(ai[j]);
        if (lane == 0) {
          red[gr * NCMAX * ES + (gc * cg + j) * ES] = sr;
          if (CPLX) red[gr * NCMAX * ES + (gc * cg + j) * ES + 1] = si;
        }
      }
    }
    if (MODE == 1 && norm_trick) {
      const double sn = warp_sum(nrm);
      if (lane == 0 && gc == 0) s_nrm[gr] = sn;
    }
  } else {
    const double sn = warp_sum(nrm);
    if (lane == 0 && gc == 0) red[gr] = sn;
  }
  asm volatile("bar.sync 9, %0;" ::"r"(kConsumerWarps * 32) : "memory");
  const int nbase = (MODE <= 1) ? ncols * ES : 1;
  const int nout = nbase + ((MODE == 1 && norm_trick) ? 1 : 0);  // values this launch reduces over the grid
  for (int t = threadIdx.x; t < nred; t += kConsumerWarps * 32) {
    double sacc = 0.0;
#pragma unroll
    for (int g = 0; g < WR; ++g) sacc += (MODE <= 1) ? red[g * NCMAX * ES + t] : red[g];
    if (t < nbase) partial[size_t(blockIdx.x) * kPartialStride + t] = sacc;
  }
  if (MODE == 1 && norm_trick && threadIdx.x == 0) {
    double sacc = 0.0;
#pragma unroll
    for (int g = 0; g < WR; ++g) sacc += s_nrm[g];
    partial[size_t(blockIdx.x) * kPartialStride + nbase] = sacc;
  }
  // ===== last CTA sums the per-CTA partials in CTA order =====
  __threadfence();
  asm volatile("bar.sync 9, %0;" ::"r"(kConsumerWarps * 32) : "memory");
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(ticket, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  asm volatile("bar.sync 9, %0;" ::"r"(kConsumerWarps * 32) : "memory");
  if (s_last) {
    __threadfence();
    // Sum the per-CTA partials in a fixed pattern (deterministic).  A thread takes a PAIR of neighbouring values (one
    // 16-byte load per CTA row) and the interleaved CTAs of its group with eight independent accumulators: ~8 G loads
    // in flight instead of one serial chain of gridDim.x L2 round trips.  Then one thread per value adds the G group
    // sums in order.
    const int tid = threadIdx.x, nb = int(gridDim.x);
    const int npair = (nout + 1) >> 1;
    int G = (kConsumerWarps * 32) / npair;
    G = G < 1 ? 1 : (G > 16 ? 16 : G);
    if (tid < G * npair) {
      const int vp = tid % npair, g = tid / npair;
      const double2* pp = reinterpret_cast<const double2*>(partial) + vp;
      constexpr int kRow = kPartialStride / 2;  // row length in double2
      double2 acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = make_double2(0.0, 0.0);
      int b = g;
      for (; b + 7 * G < nb; b += 8 * G) {
        double2 ld[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ld[u] = __ldcg(pp + size_t(b + u * G) * kRow);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc[u].x += ld[u].x;
          acc[u].y += ld[u].y;
        }
      }
      for (; b < nb; b += G) {  // fewer than 8 rows left
        const double2 l = __ldcg(pp + size_t(b) * kRow);
        acc[0].x += l.x;
        acc[0].y += l.y;
      }
      const double sx = ((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x)) + ((acc[4].x + acc[5].x) + (acc[6].x + acc[7].x));
      const double sy = ((acc[0].y + acc[1].y) + (acc[2].y + acc[3].y)) + ((acc[4].y + acc[5].y) + (acc[6].y + acc[7].y));
      s_fin[(g * npair + vp) * 2] = sx;
      s_fin[(g * npair + vp) * 2 + 1] = sy;
    }
    asm volatile("bar.sync 9, %0;" ::"r"(kConsumerWarps * 32) : "memory");
    for (int t = tid; t < nout; t += kConsumerWarps * 32) {
      double sacc = 0.0;
      for (int g = 0; g < G; ++g) sacc += s_fin[g * npair * 2 + t];
      if (push.P > 1)
        mail_push_value(push, t, sacc);  // this rank's partial straight into every peer's mailbox (NVLink)
      else
        hout[t] = sacc;
    }
    if (threadIdx.x == 0) *ticket = 0u;
    if (push.P > 1) {
      __threadfence_system();
      asm volatile("bar.sync 9, %0;" ::"r"(kConsumerWarps * 32) : "memory");
      mail_publish(push, threadIdx.x);
    }
  }
}

// Stand-alone slab push: the stores of the fused version without a Gram-Schmidt pass around them.
__global__ void __launch_bounds__(256)
slab_push_kernel(const double* __restrict__ w, const __grid_constant__ SlabPush slab, const int* __restrict__ halt) {
  if (*halt) return;
  const long long stride = (long long)gridDim.x * blockDim.x * 2;
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; e < slab.nd; e += stride) {
    const double2 yv = *reinterpret_cast<const double2*>(w + e);
    for (int d = 0; d < slab.n; ++d) {
      if (slab.kind[d] == 0) {
        if (e >= slab.lo[d] && e < slab.hi[d]) *reinterpret_cast<double2*>(slab.dst[d] + (e - slab.lo[d])) = yv;
      } else if (slab.es == 2) {
        const long long idx = e >> 1;
        if ((idx & 1) == slab.lo[d]) *reinterpret_cast<double2*>(slab.dst[d] + (idx >> 1) * 2) = yv;
      } else {
        slab.dst[d][e >> 1] = slab.lo[d] ? yv.y : yv.x;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(slab.ticket, 1u);
    if (prev == gridDim.x - 1) {
      *slab.ticket = 0u;
      __threadfence_system();
      for (int d = 0; d < slab.n; ++d) st_release_sys_u64(slab.flag[d], slab.seq);
    }
  }
}

int slab_push(cmb_ctx* ctx, const double* w, const SlabPush& slab, const int* halt) {
  if (slab.n <= 0) return CMB_OK;
  LaunchScope ls(ctx, "slab_push");
  const int grid = int(std::max<long long>(1, std::min<long long>((slab.nd / 2 + 255) / 256, (long long)ctx->num_sms * 8)));
  slab_push_kernel<<<grid, 256, 0, ctx->stream>>>(w, slab, halt);
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

// ---- host side ---------------------------------------------------------------------------------------
template <int CG, int WC, bool CPLX, int MODE>
static int launch_cfg(cmb_ctx* ctx, const CgsPass& a, int cg) {
  using Cfg = CgsCfg<WC>;
  const int64_t ntiles64 = a.ld / Cfg::T;
  CMB_REQUIRE(a.ld % Cfg::T == 0 && ntiles64 < (int64_t(1) << 31) / Cfg::T, "padded length not tileable");
  const int ntiles = int(ntiles64);
  const int nc = cg * WC;
  const int stage_bytes = (nc * Cfg::T + Cfg::T) * 8;
  const int fixed = Cfg::PW_BYTES + Cfg::WR * CG * WC * 2 * 8 + 2 * 8 * 16 + 64;
  int stages = (200 * 1024 - fixed) / stage_bytes;
  if (stages > 12) stages = 12;
  if (stages < 2) stages = 2;
  const size_t smem = size_t(stages) * stage_bytes + fixed;

  // tensor map over the chunk of V: dim0 = rows (doubles, contiguous), dim1 = columns; box = [BOXR x nc]
  CUtensorMap tm;
  cuuint64_t gdim[2] = {cuuint64_t(a.ld), cuuint64_t(a.ncols)};
  cuuint64_t gstr[1] = {cuuint64_t(a.col_stride) * 8};
  cuuint32_t box[2] = {cuuint32_t(Cfg::BOXR), cuuint32_t(nc)};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = get_encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(a.V), gdim, gstr, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) ld=%lld ncols=%d", int(cr), (long long)a.ld, a.ncols);
    return CMB_ERR_CUDA;
  }
  auto kern = cgs_kernel<CG, WC, CPLX, MODE>;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    CMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set[ctx->device & 63] = true;
  }
  int grid = ctx->num_sms;
  if (grid > ntiles) grid = ntiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  static const char* fam[3] = {"cgs_dot", "cgs_update_dot", "cgs_update_norm"};
  {
    LaunchScope ls(ctx, a.family ? a.family : fam[MODE]);
    CMB_CUDA(launch_pdl(pdl_wanted(false, ctx->nranks), kern, grid, kThreads, smem, ctx->stream, tm, a.x, a.y, a.hin, a.hout, ctx->d_partial, ctx->d_ticket,
                        a.halt, a.ncols, cg, ntiles, stages, a.pull, a.push, a.norm_trick, a.retry, a.retry_tag,
                        a.norm_guard, a.slab, a.v_stable ? 1 : 0));
  }
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

template <bool CPLX, int MODE>
static int launch_mode(cmb_ctx* ctx, const CgsPass& a) {
  const int c = a.ncols;
  // Fewest column groups WC such that a warp's share fits the register budget: fewer groups mean taller tiles
  // (T = 64 * 8 / WC rows) and less (for WC = 1: no) cross-warp exchange per tile.
  if (c <= 8) return launch_cfg<8, 1, CPLX, MODE>(ctx, a, c);
  if constexpr (!CPLX) {
    if (c <= 16) return launch_cfg<16, 1, CPLX, MODE>(ctx, a, c);
    if (c <= 32) return launch_cfg<16, 2, CPLX, MODE>(ctx, a, (c + 1) / 2);
    if (c <= 64) return launch_cfg<16, 4, CPLX, MODE>(ctx, a, (c + 3) / 4);
    if (c <= 128) return launch_cfg<16, 8, CPLX, MODE>(ctx, a, (c + 7) / 8);
  } else {
    if (c <= 16) return launch_cfg<8, 2, CPLX, MODE>(ctx, a, (c + 1) / 2);
    if (c <= 32) return launch_cfg<8, 4, CPLX, MODE>(ctx, a, (c + 3) / 4);
    if (c <= 64) return launch_cfg<8, 8, CPLX, MODE>(ctx, a, (c + 7) / 8);
  }
  set_error("cgs pass: %d columns exceed the per-pass maximum", c);
  return CMB_ERR_INVALID;
}

int cgs_pass(cmb_ctx* ctx, bool cplx, int mode, const CgsPass& a) {
  CMB_REQUIRE(a.ncols >= 1 && a.ncols <= cgs_max_cols(cplx), "bad column count");
  CMB_REQUIRE(mode >= 0 && mode <= 2, "bad mode");
  if (cplx) {
    if (mode == 0) return launch_mode<true, 0>(ctx, a);
    if (mode == 1) return launch_mode<true, 1>(ctx, a);
    return launch_mode<true, 2>(ctx, a);
  }
  if (mode == 0) return launch_mode<false, 0>(ctx, a);
  if (mode == 1) return launch_mode<false, 1>(ctx, a);
  return launch_mode<false, 2>(ctx, a);
}

}  // namespace cmb
