// compose.cu — device-resident operator algebra: linear combinations and products of
// operators that already live in HBM, as operators themselves.
//
// The reference builds sums and products of operators on the host (vector_map.hpp:38-266: every application runs the
// wrapped std::functions one after the other and adds the results).  Here the same algebra stays on the device: a
// LinearOp applies its terms and accumulates with fused axpy kernels, a ProductOp chains two applies through a scratch
// vector; neither copies a vector to the host.  The first child apply carries the step prologue of the fused operator
// interface (u = w / beta, breakdown test); the others run on the already normalised u with beta = 1.
#include <algorithm>
#include <vector>

#include "op.cuh"

namespace cmb {

template <bool CPLX>
__global__ void __launch_bounds__(256) scale_inplace_kernel(double* __restrict__ v, double cr, double ci, long long n2,
                                                            const int* __restrict__ halt) {
  if (*halt) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 a = reinterpret_cast<double2*>(v)[i];
    if (CPLX)
      a = make_double2(cr * a.x - ci * a.y, cr * a.y + ci * a.x);
    else
      a = make_double2(cr * a.x, cr * a.y);
    reinterpret_cast<double2*>(v)[i] = a;
  }
}

struct ComposeBase : cmb_op {
  double *t_u = nullptr, *t_v = nullptr, *scal = nullptr;  // scratch vectors (ld doubles) and scalars
  int64_t ld = 0;
  ~ComposeBase() override {
    if (!ctx) return;
    pool_free(ctx, t_u);
    pool_free(ctx, t_v);
    pool_free(ctx, scal);
  }
  int alloc() {
    const int64_t nd = n_local * (cplx ? 2 : 1);
    ld = (nd + 511) / 512 * 512;
    CMB_TRY(pool_alloc(ctx, &t_u, sizeof(double) * size_t(ld)));
    CMB_TRY(pool_alloc(ctx, &t_v, sizeof(double) * size_t(ld)));
    CMB_TRY(pool_alloc(ctx, &scal, sizeof(double) * 8));
    CMB_CUDA(cudaMemsetAsync(t_u, 0, sizeof(double) * size_t(ld), ctx->stream));
    CMB_CUDA(cudaMemsetAsync(t_v, 0, sizeof(double) * size_t(ld), ctx->stream));
    const double init[8] = {1.0, 0, 0, 0, 0, 0, 0, 0};  // scal[0] = 1: "norm^2" of an already normalised input
    CMB_CUDA(cudaMemcpyAsync(scal, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return CMB_OK;
  }
  // step scalars for a child that runs on an already normalised vector
  StepScalars plain(const StepScalars& sc) const {
    StepScalars s;
    s.nrm2 = scal;
    s.threshold = -1.0;
    s.halt = sc.halt;
    s.beta_slot = scal + 2;
    s.alpha_slot = scal + 4;
    return s;
  }
  int scale(double* v, double cr, double ci, const int* halt) {
    if (cr == 1.0 && ci == 0.0) return CMB_OK;
    LaunchScope ls(ctx, "vec_scale");
    const long long n2 = ld / 2;
    const int grid = int(std::max<long long>(1, std::min<long long>((n2 + 255) / 256, (long long)ctx->num_sms * 8)));
    if (cplx)
      scale_inplace_kernel<true><<<grid, 256, 0, ctx->stream>>>(v, cr, ci, n2, halt);
    else
      scale_inplace_kernel<false><<<grid, 256, 0, ctx->stream>>>(v, cr, ci, n2, halt);
    CMB_CUDA(cudaGetLastError());
    return CMB_OK;
  }
  // v += shift u ; alpha = <u, v>
  int finish(double* ucol, double* v, double shr, double shi, const StepScalars& sc) {
    if (shr != 0.0 || shi != 0.0) CMB_TRY(vec_axpy_shift(ctx, cplx, shr, shi, ucol, v, ld, sc.halt));
    return vec_dot(ctx, cplx, ucol, v, ld, sc.alpha_slot, sc.halt);
  }
};

struct LinearOp : ComposeBase {
  std::vector<cmb_op*> terms;
  std::vector<double> cr, ci;
  int apply(const double* w, double* ucol, double* v, double shr, double shi, const StepScalars& sc) override {
    StepScalars first = sc;
    first.alpha_slot = scal + 6;  // the children's own alpha dots are not used: alpha is taken from the sum
    CMB_TRY(terms[0]->apply(w, ucol, v, 0.0, 0.0, first));
    CMB_TRY(scale(v, cr[0], ci[0], sc.halt));
    const StepScalars s2 = plain(sc);
    for (size_t i = 1; i < terms.size(); ++i) {
      CMB_TRY(terms[i]->apply(ucol, t_u, t_v, 0.0, 0.0, s2));
      CMB_TRY(vec_axpy_shift(ctx, cplx, cr[i], ci[i], t_v, v, ld, sc.halt));
    }
    return finish(ucol, v, shr, shi, sc);
  }
};

struct ProductOp : ComposeBase {
  cmb_op *outer = nullptr, *inner = nullptr;
  int apply(const double* w, double* ucol, double* v, double shr, double shi, const StepScalars& sc) override {
    StepScalars first = sc;
    first.alpha_slot = scal + 6;
    CMB_TRY(inner->apply(w, ucol, t_v, 0.0, 0.0, first));      // u = w / beta ; t_v = B u
    CMB_TRY(outer->apply(t_v, t_u, v, 0.0, 0.0, plain(sc)));   // v = A (B u)
    return finish(ucol, v, shr, shi, sc);
  }
};

}  // namespace cmb

using namespace cmb;

static int check_children(cmb_ctx* ctx, cmb_op* const* ops, int64_t n) {
  CMB_REQUIRE(ctx && ops && n >= 1, "bad argument");
  for (int64_t i = 0; i < n; ++i) {
    CMB_REQUIRE(ops[i], "null operator");
    CMB_REQUIRE(ops[i]->ctx == ctx, "operators of a composition must live on the same context");
    CMB_REQUIRE(ops[i]->dtype == ops[0]->dtype && ops[i]->n_global == ops[0]->n_global && ops[i]->n_local == ops[0]->n_local &&
                    ops[i]->row_begin == ops[0]->row_begin,
                "operators of a composition must have the same dtype, shape and row range");
  }
  return CMB_OK;
}

static void inherit(cmb_op* op, const cmb_op* from, const char* family) {
  op->ctx = from->ctx;
  op->dtype = from->dtype;
  op->cplx = from->cplx;
  op->n_global = from->n_global;
  op->row_begin = from->row_begin;
  op->n_local = from->n_local;
  op->family = family;
}

extern "C" {

int cmb_op_linear_create(cmb_ctx* ctx, int64_t nterms, cmb_op* const* ops, const void* coefs, cmb_op** out) {
  CMB_REQUIRE(out && coefs, "null argument");
  *out = nullptr;
  CMB_TRY(check_children(ctx, ops, nterms));
  CMB_CUDA(cudaSetDevice(ctx->device));
  LinearOp* op = new (std::nothrow) LinearOp();
  if (!op) return CMB_ERR_NOMEM;
  inherit(op, ops[0], "op_linear");
  const double* c = static_cast<const double*>(coefs);
  const double s = op->cplx ? 16.0 : 8.0;
  for (int64_t i = 0; i < nterms; ++i) {
    op->terms.push_back(ops[i]);
    op->cr.push_back(op->cplx ? c[2 * i] : c[i]);
    op->ci.push_back(op->cplx ? c[2 * i + 1] : 0.0);
    op->bytes += ops[i]->bytes + (i ? 3.0 : 2.0) * double(op->n_local) * s;
  }
  int rc = op->alloc();
  if (rc != CMB_OK) {
    delete op;
    return rc;
  }
  *out = op;
  return CMB_OK;
}

int cmb_op_product_create(cmb_ctx* ctx, cmb_op* outer, cmb_op* inner, cmb_op** out) {
  CMB_REQUIRE(out, "null argument");
  *out = nullptr;
  cmb_op* both[2] = {outer, inner};
  CMB_TRY(check_children(ctx, both, 2));
  CMB_CUDA(cudaSetDevice(ctx->device));
  ProductOp* op = new (std::nothrow) ProductOp();
  if (!op) return CMB_ERR_NOMEM;
  inherit(op, inner, "op_product");
  op->outer = outer;
  op->inner = inner;
  op->bytes = outer->bytes + inner->bytes + 2.0 * double(op->n_local) * (op->cplx ? 16.0 : 8.0);
  int rc = op->alloc();
  if (rc != CMB_OK) {
    delete op;
    return rc;
  }
  *out = op;
  return CMB_OK;
}

}  // extern "C"
