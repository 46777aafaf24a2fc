// vecops.cu — small fused level-1 kernels (K6/K8 of SURVEY.md §2a): dot/norm, scaling, shift-axpy, phase fix.
// Used for the start vector (lanczos.hpp:299-323), the callback-operator path and Ritz-vector finishing
// (lanczos.hpp:805-816).  All HBM-bound streaming kernels: 16-byte loads, grid = a multiple of the SM count.
#include <algorithm>

#include "device_utils.cuh"

namespace cmb {

static inline int stream_grid(cmb_ctx* ctx, long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  return int(std::max<long long>(1, std::min<long long>(blocks, (long long)ctx->num_sms * 8)));
}

template <bool CPLX>
__global__ void __launch_bounds__(256)
dot_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n2 /* double2 count */,
           double* __restrict__ out, const int* __restrict__ halt, double* partial, unsigned* ticket) {
  if (*halt) return;
  const double2* a2 = reinterpret_cast<const double2*>(a);
  const double2* b2 = reinterpret_cast<const double2*>(b);
  const long long stride = (long long)gridDim.x * blockDim.x;
  double d0 = 0.0, d1 = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const double2 p = a2[i], q = b2[i];
    d0 = fma(p.x, q.x, d0);
    d0 = fma(p.y, q.y, d0);
    if (CPLX) {
      d1 = fma(p.x, q.y, d1);
      d1 = fma(-p.y, q.x, d1);
    }
  }
  grid_sum_finalize<CPLX ? 2 : 1>(d0, d1, partial, ticket, out);
}

int vec_dot(cmb_ctx* ctx, bool cplx, const double* a, const double* b, int64_t ld, double* out, const int* halt) {
  const long long n2 = ld / 2;
  const int grid = stream_grid(ctx, n2, 256 * 4);
  LaunchScope ls(ctx, "vec_dot");
  if (cplx)
    dot_kernel<true><<<grid, 256, 0, ctx->stream>>>(a, b, n2, out, halt, ctx->d_partial, ctx->d_ticket + 2);
  else
    dot_kernel<false><<<grid, 256, 0, ctx->stream>>>(a, b, n2, out, halt, ctx->d_partial, ctx->d_ticket + 2);
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

__global__ void __launch_bounds__(256)
scale_rsqrt_kernel(const double* __restrict__ x, const double* __restrict__ nrm2, double* __restrict__ y, long long n2,
                   const int* __restrict__ halt) {
  if (*halt) return;
  const double inv = 1.0 / sqrt(*nrm2);
  const double2* x2 = reinterpret_cast<const double2*>(x);
  double2* y2 = reinterpret_cast<double2*>(y);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const double2 p = x2[i];
    y2[i] = make_double2(p.x * inv, p.y * inv);
  }
}

int vec_scale_rsqrt(cmb_ctx* ctx, const double* x, const double* nrm2, double* y, int64_t ld, const int* halt) {
  const long long n2 = ld / 2;
  const int grid = stream_grid(ctx, n2, 256 * 4);
  LaunchScope ls(ctx, "vec_scale");
  scale_rsqrt_kernel<<<grid, 256, 0, ctx->stream>>>(x, nrm2, y, n2, halt);
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

template <bool CPLX>
__global__ void __launch_bounds__(256)
axpy_shift_kernel(double sr, double si, const double* __restrict__ x, double* __restrict__ y, long long n2,
                  const int* __restrict__ halt) {
  if (*halt) return;
  const double2* x2 = reinterpret_cast<const double2*>(x);
  double2* y2 = reinterpret_cast<double2*>(y);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const double2 p = x2[i];
    double2 q = y2[i];
    if (CPLX) {
      q.x += sr * p.x - si * p.y;
      q.y += sr * p.y + si * p.x;
    } else {
      q.x = fma(sr, p.x, q.x);
      q.y = fma(sr, p.y, q.y);
    }
    y2[i] = q;
  }
}

int vec_axpy_shift(cmb_ctx* ctx, bool cplx, double sr, double si, const double* x, double* y, int64_t ld,
                   const int* halt) {
  const long long n2 = (ld + 1) / 2;
  const int grid = stream_grid(ctx, n2, 256 * 4);
  LaunchScope ls(ctx, "vec_axpy");
  if (cplx)
    axpy_shift_kernel<true><<<grid, 256, 0, ctx->stream>>>(sr, si, x, y, n2, halt);
  else
    axpy_shift_kernel<false><<<grid, 256, 0, ctx->stream>>>(sr, si, x, y, n2, halt);
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

// x *= conj(phase)/|phase| / sqrt(nrm2), phase = *phase_src (the first non-zero element, lanczos.hpp:806-816)
template <bool CPLX>
__global__ void __launch_bounds__(256)
scale_phase_kernel(double* __restrict__ x, const double* __restrict__ nrm2, const double* __restrict__ phase_src,
                   long long n2) {
  const double nrm = sqrt(*nrm2);
  const double inv = nrm > 0.0 ? 1.0 / nrm : 1.0;
  double fr = 1.0, fi = 0.0;  // 1/phase = conj(phase) for |phase| = 1
  if (phase_src) {
    if (CPLX) {
      const double pr = phase_src[0], pi = phase_src[1];
      const double a = sqrt(pr * pr + pi * pi);
      if (a > 0.0) {
        fr = pr / a;
        fi = -pi / a;
      }
    } else {
      const double p = phase_src[0];
      if (p < 0.0) fr = -1.0;
    }
  }
  fr *= inv;
  fi *= inv;
  double2* x2 = reinterpret_cast<double2*>(x);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 p = x2[i];
    if (CPLX)
      p = make_double2(p.x * fr - p.y * fi, p.x * fi + p.y * fr);
    else
      p = make_double2(p.x * fr, p.y * fr);
    x2[i] = p;
  }
}

int vec_scale_phase(cmb_ctx* ctx, bool cplx, double* x, const double* nrm2, const double* phase_src, int64_t ld) {
  const long long n2 = ld / 2;
  const int grid = stream_grid(ctx, n2, 256 * 4);
  LaunchScope ls(ctx, "vec_scale");
  if (cplx)
    scale_phase_kernel<true><<<grid, 256, 0, ctx->stream>>>(x, nrm2, phase_src, n2);
  else
    scale_phase_kernel<false><<<grid, 256, 0, ctx->stream>>>(x, nrm2, phase_src, n2);
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

__global__ void __launch_bounds__(256)
real_to_complex_kernel(const double* __restrict__ x, double* __restrict__ z, long long n) {
  double2* z2 = reinterpret_cast<double2*>(z);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    z2[i] = make_double2(x[i], 0.0);
}

int vec_real_to_complex(cmb_ctx* ctx, const double* x, double* z, int64_t n) {
  const int grid = stream_grid(ctx, n, 256 * 4);
  LaunchScope ls(ctx, "vec_scale");
  real_to_complex_kernel<<<grid, 256, 0, ctx->stream>>>(x, z, n);
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

template <bool CPLX>
__global__ void __launch_bounds__(256)
first_nonzero_kernel(const double* __restrict__ x, long long n, long long offset, unsigned long long* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if ((unsigned long long)(i + offset) >= *reinterpret_cast<volatile unsigned long long*>(out)) return;  // an earlier index won
    bool nz;
    if (CPLX)
      nz = (x[2 * i] != 0.0) || (x[2 * i + 1] != 0.0);
    else
      nz = x[i] != 0.0;
    if (nz) {
      atomicMin(out, (unsigned long long)(i + offset));
      return;
    }
  }
}

// out[0] = global index (local index + offset) of the first element with |x_i| > 0, ULLONG_MAX if none
int vec_first_nonzero(cmb_ctx* ctx, bool cplx, const double* x, int64_t n, int64_t offset, unsigned long long* out) {
  CMB_CUDA(cudaMemsetAsync(out, 0xff, sizeof(unsigned long long), ctx->stream));
  if (n == 0) return CMB_OK;
  const int grid = stream_grid(ctx, n, 256);
  LaunchScope ls(ctx, "vec_first_nonzero");
  if (cplx)
    first_nonzero_kernel<true><<<grid, 256, 0, ctx->stream>>>(x, n, offset, out);
  else
    first_nonzero_kernel<false><<<grid, 256, 0, ctx->stream>>>(x, n, offset, out);
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

// out[0..es) = x[idx - offset] if the global index idx[0] falls into this rank's rows, else 0
__global__ void pick_element_kernel(const double* __restrict__ x, const unsigned long long* __restrict__ idx,
                                    long long offset, long long n, int es, double* __restrict__ out) {
  const unsigned long long g = idx[0];
  for (int e = 0; e < es; ++e) {
    double v = 0.0;
    if (g != ~0ull && (long long)g >= offset && (long long)g < offset + n) v = x[((long long)g - offset) * es + e];
    out[e] = v;
  }
}

int vec_pick_element(cmb_ctx* ctx, const double* x, const unsigned long long* idx, int64_t offset, int64_t n, int es,
                     double* out) {
  LaunchScope ls(ctx, "vec_first_nonzero");
  pick_element_kernel<<<1, 1, 0, ctx->stream>>>(x, idx, offset, n, es, out);
  CMB_CUDA(cudaGetLastError());
  return CMB_OK;
}

}  // namespace cmb
