// krylov.cu — device-resident Krylov state and the step functions of the C-ABI.
//
// cmb_krylov holds what LanczosBase keeps in lanczosvectors_/v_/alpha_/beta_ (lanczos.hpp:233-239) and
// ArnoldiBase in arnoldivectors_/v_/h_/residue_ (arnoldi.hpp:181-187), in HBM:
//   basis    column segments of `seg_cols` columns, each column `ld` doubles (ld = local length padded to a
//            multiple of 512, pads zero; a complex column is ld/2 interleaved pairs).  Deflation vectors
//            (orthogonalizingVectors_) occupy the first ndefl columns, Krylov vectors follow.
//   v        operator output (A+shift) u_last
//   w        the next, not yet normalised, basis vector; scal[0] = ||w||^2
// One Lanczos step = [CGS2 of v against the basis: three fused passes] + [operator apply fused with
// w/beta -> new column and the alpha dot].  One Arnoldi step = the same two pieces in the other order.
// Scalars stay on the device (kernels read beta = sqrt(scal[0]) themselves and stop the chain through a
// sticky halt flag), so cmb_lanczos_run() can enqueue many steps with no host round trip.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "cmpt_b200_debug.h"
#include "op.cuh"

using namespace cmb;

namespace {
constexpr int kMaxSlots = 1 << 16;
struct Chunk {
  const double* V;
  int ncols;
  int64_t col_stride;
};
}  // namespace

struct cmb_krylov {
  cmb_ctx* ctx = nullptr;
  int dtype = CMB_F64;
  bool cplx = false;
  int es = 1;
  int64_t n_global = 0, row_begin = 0, n_local = 0, nd_local = 0, ld = 0;
  int seg_cols = 0;
  std::vector<double*> segs;
  int ndefl = 0, nk = 0;
  double *v = nullptr, *w = nullptr;
  double *h1 = nullptr, *h2 = nullptr;
  int hcap = 0;
  double* scal = nullptr;  // [0] ||w||^2, [1] scratch norm, [2..3] scratch, [4..5] scratch alpha
  int* halt = nullptr;
  double *alpha_dev = nullptr, *beta_dev = nullptr;
  double* h_stage = nullptr;  // pinned staging for scalars
  size_t h_stage_elems = 0;
  bool started = false;
  unsigned long long nrm2_seq = 0;  // non-zero: ||w||^2 is the mailbox reduction with this sequence number
  double* d_start = nullptr;        // the start vector of the last cmb_krylov_start (for cmb_krylov_restart)
  bool w_pushed = false;            // the pass that wrote w also pushed it to the partner ranks (SlabPush)
  bool resume_norm_ready = false;   // w is orthogonalised and scal[0] holds its explicitly reduced norm (guard retry)
  double residue = 0.0;  // Arnoldi: last residual norm (host copy)
  double bytes = 0.0;
  double *tmp1 = nullptr, *tmp2 = nullptr, *tmpz = nullptr;
  // Ritz-vector assembly: two output buffers so that the device->host copy of one vector (copy stream) overlaps the
  // assembly of the next
  double* xout[2] = {nullptr, nullptr};
  size_t xout_doubles = 0;
  cudaEvent_t ev_x[2] = {nullptr, nullptr}, ev_c[2] = {nullptr, nullptr};
  unsigned long long* d_idx = nullptr;

  double* col(int j) const { return segs[j / seg_cols] + int64_t(j % seg_cols) * ld; }
};

static int ensure_cols(cmb_krylov* K, int total) {
  while (int(K->segs.size()) * K->seg_cols < total) {
    double* p = nullptr;
    const size_t bytes = sizeof(double) * size_t(K->seg_cols) * size_t(K->ld);
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      set_error("out of device memory growing the Krylov basis to %d columns (%.2f GB per %d-column segment)", total,
                double(bytes) / 1e9, K->seg_cols);
      return CMB_ERR_NOMEM;
    }
    CMB_CUDA(cudaMemsetAsync(p, 0, bytes, K->ctx->stream));
    K->segs.push_back(p);
  }
  const int cap = int(K->segs.size()) * K->seg_cols;
  if (cap > K->hcap) {
    // stream-ordered (re)allocation: cudaFree would synchronise the whole device in the middle of a step chain,
    // which dead-locks virtual ranks (a peer's kernel may be spinning on this rank's next launch)
    pool_free(K->ctx, K->h1);
    pool_free(K->ctx, K->h2);
    K->h1 = K->h2 = nullptr;
    CMB_TRY(pool_alloc(K->ctx, &K->h1, sizeof(double) * 2 * (cap + 8)));
    CMB_TRY(pool_alloc(K->ctx, &K->h2, sizeof(double) * 2 * (cap + 8)));
    K->hcap = cap;
  }
  return CMB_OK;
}

static int ensure_stage(cmb_krylov* K, size_t elems) {
  if (elems <= K->h_stage_elems) return CMB_OK;
  hfree(K->ctx, K->h_stage);
  K->h_stage = nullptr;
  K->h_stage_elems = 0;
  size_t want = std::max<size_t>(elems * 2, 4096);
  CMB_CUDA(cudaMallocHost(&K->h_stage, sizeof(double) * want));
  K->h_stage_elems = want;
  return CMB_OK;
}

// contiguous chunks covering absolute columns [c0, c1)
static void contiguous_chunks(const cmb_krylov* K, int c0, int c1, std::vector<Chunk>& out) {
  const int maxc = std::min(K->seg_cols, cgs_max_cols(K->cplx));
  int c = c0;
  while (c < c1) {
    const int in_seg = K->seg_cols - (c % K->seg_cols);
    const int n = std::min(std::min(in_seg, maxc), c1 - c);
    out.push_back({K->col(c), n, K->ld});
    c += n;
  }
}

// chunks covering the strided set {first, first+stride, ...} < end (absolute column indices)
static void strided_chunks(const cmb_krylov* K, int first, int end, int stride, std::vector<Chunk>& out) {
  if (stride == 1) {
    contiguous_chunks(K, first, end, out);
    return;
  }
  const int maxc = cgs_max_cols(K->cplx);
  int c = first;
  while (c < end) {
    const int seg = c / K->seg_cols;
    int n = 0;
    int cc = c;
    while (cc < end && cc / K->seg_cols == seg && n < maxc) {
      ++n;
      cc += stride;
    }
    out.push_back({K->col(c), n, int64_t(stride) * K->ld});
    c = cc;
  }
}

static int total_cols(const std::vector<Chunk>& ch) {
  int t = 0;
  for (auto& c : ch) t += c.ncols;
  return t;
}

// Gram-Schmidt of x against the chunk columns, result in y (may alias x); coefficients of the first pass
// in h1, of the second in h2 (chunk order); ||y||^2 in nrm2_out.  Single chunk: DOT, UPDATE_DOT,
// UPDATE_NORM (the basis is streamed three times).  Several chunks: one extra DOT sweep.
// Fused variant for row-partitioned runs (one chunk): the three passes exchange their per-rank partial results
// through the peer-memory mailboxes (NVLink stores from the finishing CTA, summed in the next kernel's prologue),
// so no collective kernel runs in between.  *nrm2_seq receives the sequence number under which ||y||^2 will be
// found by the consumer (the operator apply).
static int gram_schmidt2_mailed(cmb_krylov* K, const Chunk& c, const double* x, double* y,
                                unsigned long long* nrm2_seq, int retry_tag, cmb_op* push_op = nullptr) {
  cmb_ctx* ctx = K->ctx;
  CgsPass p;
  p.ld = K->ld;
  p.halt = K->halt;
  p.V = c.V;
  p.ncols = c.ncols;
  p.col_stride = c.col_stride;
  // pass 1: partial h1 -> mailboxes
  p.x = x;
  p.push = mail_next_push(ctx);
  const unsigned long long s1 = p.push.seq;
  CMB_TRY(cgs_pass(ctx, K->cplx, CGS_DOT, p));
  p.v_stable = true;  // the passes below follow a pass over the same columns
  // pass 2: h1 = sum of partials ; y = x - V h1 ; partial h2 -> mailboxes
  // (also ||y||^2, so that the third pass needs no reduction of its own: ||y - V h2||^2 = ||y||^2 - |h2|^2)
  // The identity needs orthonormal columns: the Krylov vectors are (to rounding), user-supplied deflation vectors
  // need not be, so with deflation vectors the norm keeps its own reduction.
  const bool trick = (K->ndefl == 0);
  p.y = y;
  p.pull = mail_pull_of(ctx, s1, K->h1);
  p.pull.halt = K->halt;
  p.push = mail_next_push(ctx);
  p.norm_trick = trick ? 1 : 0;
  const unsigned long long s2 = p.push.seq;
  CMB_TRY(cgs_pass(ctx, K->cplx, CGS_UPDATE_DOT, p));
  p.x = y;
  p.pull = mail_pull_of(ctx, s2, K->h2);
  p.pull.halt = K->halt;
  if (trick) {
    // pass 3: y -= V h2 ; ||y||^2 from the reduced h2 and ||y_before||^2 (same bits on every rank) -> K->scal[0]
    p.push = MailPush();
    p.hout = K->scal;
    p.retry = K->halt + 1;  // cancellation guard: halt[1] = 1, halt[2] = retry_tag, and the chain halts
    p.retry_tag = retry_tag;
    p.norm_guard = ctx->norm_guard;
    *nrm2_seq = 0;  // the norm is a plain device scalar: the operator apply needs no mailbox
  } else {
    // pass 3: y -= V h2 ; partial ||y||^2 -> mailboxes, summed in the prologue of the operator apply
    p.push = mail_next_push(ctx);
    *nrm2_seq = p.push.seq;
  }
  if (push_op && y == K->w && push_op->slab_push_begin(y, &p.slab)) K->w_pushed = true;  // this pass also pushes y to the partners
  CMB_TRY(cgs_pass(ctx, K->cplx, CGS_UPDATE_NORM, p));
  return CMB_OK;
}

static int gram_schmidt2(cmb_krylov* K, const std::vector<Chunk>& chunks, const double* x, double* y,
                         double* nrm2_out, cmb_op* push_op = nullptr) {
  cmb_ctx* ctx = K->ctx;
  const int es = K->es;
  const int nch = int(chunks.size());
  const int ctot = total_cols(chunks);
  CgsPass p;
  p.ld = K->ld;
  p.halt = K->halt;
  auto setc = [&](const Chunk& c) {
    p.V = c.V;
    p.ncols = c.ncols;
    p.col_stride = c.col_stride;
  };
  // pass 1: h1 = V^H x
  int off = 0;
  for (int i = 0; i < nch; ++i) {
    setc(chunks[i]);
    p.x = x;
    p.y = nullptr;
    p.hin = nullptr;
    p.hout = K->h1 + off * es;
    CMB_TRY(cgs_pass(ctx, K->cplx, CGS_DOT, p));
    p.v_stable = true;  // every later pass follows a pass (or a reduction) that leaves the basis alone
    off += chunks[i].ncols;
  }
  CMB_TRY(allreduce_sum_f64(ctx, K->h1, size_t(ctot) * es));
  // pass 2: y = x - V h1 ; h2 = V^H y
  off = 0;
  for (int i = 0; i < nch; ++i) {
    setc(chunks[i]);
    p.x = (i == 0) ? x : y;
    p.y = y;
    p.hin = K->h1 + off * es;
    if (i == nch - 1) {
      p.hout = K->h2 + off * es;
      CMB_TRY(cgs_pass(ctx, K->cplx, CGS_UPDATE_DOT, p));
    } else {
      p.hout = K->scal + 1;
      CMB_TRY(cgs_pass(ctx, K->cplx, CGS_UPDATE_NORM, p));
    }
    off += chunks[i].ncols;
  }
  off = 0;
  for (int i = 0; i < nch - 1; ++i) {
    setc(chunks[i]);
    p.x = y;
    p.y = nullptr;
    p.hin = nullptr;
    p.hout = K->h2 + off * es;
    CMB_TRY(cgs_pass(ctx, K->cplx, CGS_DOT, p));
    off += chunks[i].ncols;
  }
  CMB_TRY(allreduce_sum_f64(ctx, K->h2, size_t(ctot) * es));
  // pass 3: y -= V h2 ; ||y||^2
  off = 0;
  for (int i = 0; i < nch; ++i) {
    setc(chunks[i]);
    p.x = y;
    p.y = y;
    p.hin = K->h2 + off * es;
    p.hout = (i == nch - 1) ? nrm2_out : K->scal + 1;
    if (i == nch - 1 && push_op && y == K->w && push_op->slab_push_begin(y, &p.slab)) K->w_pushed = true;
    CMB_TRY(cgs_pass(ctx, K->cplx, CGS_UPDATE_NORM, p));
    off += chunks[i].ncols;
  }
  CMB_TRY(allreduce_sum_f64(ctx, nrm2_out, 1));
  return CMB_OK;
}

// y = x - sum_chunks V hin ; ||y||^2 -> nrm2_out (coefficients already on the device, chunk order)
static int subtract_cols(cmb_krylov* K, const std::vector<Chunk>& chunks, const double* hin, const double* x, double* y,
                         double* nrm2_out, const char* family) {
  CgsPass p;
  p.ld = K->ld;
  p.halt = K->halt;
  p.family = family;
  int off = 0;
  const int nch = int(chunks.size());
  for (int i = 0; i < nch; ++i) {
    p.V = chunks[i].V;
    p.ncols = chunks[i].ncols;
    p.col_stride = chunks[i].col_stride;
    p.x = (i == 0) ? x : y;
    p.y = y;
    p.hin = hin + off * K->es;
    p.hout = (i == nch - 1) ? nrm2_out : K->scal + 1;
    CMB_TRY(cgs_pass(K->ctx, K->cplx, CGS_UPDATE_NORM, p));
    p.v_stable = true;
    off += chunks[i].ncols;
  }
  CMB_TRY(allreduce_sum_f64(K->ctx, nrm2_out, 1));
  return CMB_OK;
}

// Operator apply on K->w.  A push announced for w is honoured only if the Krylov state says w still is what was pushed.
static int apply_w(cmb_krylov* K, cmb_op* op, double* ucol, double shr, double shi, const StepScalars& sc) {
  if (!K->w_pushed) op->slab_push_cancel();
  K->w_pushed = false;
  return op->apply(K->w, ucol, K->v, shr, shi, sc);
}

static void add_step_bytes(cmb_krylov* K, const cmb_op* op, int c) {
  // B_step(c) = B_op + (3c + 7) n s
  K->bytes += op->bytes + (3.0 * c + 7.0) * double(K->n_local) * (K->cplx ? 16.0 : 8.0);
}

static int check_pair(const cmb_krylov* K, const cmb_op* op) {
  CMB_REQUIRE(K && op, "null argument");
  CMB_REQUIRE(K->ctx == op->ctx, "operator and Krylov state belong to different contexts");
  CMB_REQUIRE(K->dtype == op->dtype, "operator and Krylov state have different dtypes");
  CMB_REQUIRE(K->n_global == op->n_global && K->n_local == op->n_local && K->row_begin == op->row_begin,
              "operator and Krylov state have different shapes");
  return CMB_OK;
}

// Enqueue the orthogonalisation part of Lanczos step k (k = index of the newest Krylov vector).
static int enqueue_lanczos_orth(cmb_krylov* K, int64_t interval, int retry_tag, cmb_op* push_op) {
  const int k = K->nk - 1;
  std::vector<Chunk> chunks;
  if (interval == 1) {
    // full reorthogonalisation: the three-term recurrence is subsumed by CGS pass 1
    // (alpha_k = h1[k], beta_{k-1} ~ h1[k-1]); deflation vectors are just the leading columns.
    contiguous_chunks(K, 0, K->ndefl + k + 1, chunks);
    K->nrm2_seq = 0;
    if (K->ctx->mail_ok && chunks.size() == 1)
      return gram_schmidt2_mailed(K, chunks[0], K->v, K->w, &K->nrm2_seq, retry_tag, push_op);
    return gram_schmidt2(K, chunks, K->v, K->w, K->scal, push_op);
  }
  K->nrm2_seq = 0;
  // explicit recurrence w = v - alpha_k u_k - beta_{k-1} u_{k-1} (lanczos.hpp:402-408)
  const int first = (k > 0) ? k - 1 : k;
  contiguous_chunks(K, K->ndefl + first, K->ndefl + k + 1, chunks);
  CMB_CUDA(cudaMemsetAsync(K->h1, 0, sizeof(double) * 4, K->ctx->stream));
  int slot = 0;
  if (k > 0) {
    CMB_CUDA(cudaMemcpyAsync(K->h1 + slot * K->es, K->beta_dev + (k - 1), sizeof(double), cudaMemcpyDeviceToDevice,
                             K->ctx->stream));
    ++slot;
  }
  CMB_CUDA(cudaMemcpyAsync(K->h1 + slot * K->es, K->alpha_dev + size_t(k) * 2, sizeof(double), cudaMemcpyDeviceToDevice,
                           K->ctx->stream));
  CMB_TRY(subtract_cols(K, chunks, K->h1, K->v, K->w, K->scal, "recurrence"));
  if (interval > 1) {
    const int kmod = int((k + 1) % interval);
    chunks.clear();
    if (kmod == 0 && K->ndefl > 0) contiguous_chunks(K, 0, K->ndefl, chunks);
    strided_chunks(K, K->ndefl + kmod, K->ndefl + k + 1, int(interval), chunks);
    if (!chunks.empty()) CMB_TRY(gram_schmidt2(K, chunks, K->w, K->w, K->scal));
  }
  return CMB_OK;
}

static int orth_cols_count(const cmb_krylov* K, int64_t interval) {
  const int k = K->nk - 1;
  if (interval == 1) return K->ndefl + k + 1;
  if (interval <= 0) return 0;
  const int kmod = int((k + 1) % interval);
  int c = (kmod == 0) ? K->ndefl : 0;
  for (int kk = kmod; kk < k + 1; kk += int(interval)) ++c;
  return c;
}

extern "C" {

int cmb_krylov_create(cmb_ctx* ctx, cmb_dtype dtype, int64_t n_global, int64_t row_begin, int64_t row_end,
                      int64_t reserve_cols, cmb_krylov** out) {
  CMB_REQUIRE(ctx && out, "null argument");
  *out = nullptr;
  CMB_REQUIRE(dtype == CMB_F64 || dtype == CMB_C64, "dtype must be CMB_F64 or CMB_C64");
  CMB_REQUIRE(n_global >= 1 && row_begin >= 0 && row_begin < row_end && row_end <= n_global, "bad row range");
  CMB_CUDA(cudaSetDevice(ctx->device));
  cmb_krylov* K = new (std::nothrow) cmb_krylov();
  if (!K) return CMB_ERR_NOMEM;
  K->ctx = ctx;
  K->dtype = dtype;
  K->cplx = dtype == CMB_C64;
  K->es = K->cplx ? 2 : 1;
  K->n_global = n_global;
  K->row_begin = row_begin;
  K->n_local = row_end - row_begin;
  K->nd_local = K->n_local * K->es;
  K->ld = (K->nd_local + 511) / 512 * 512;
  const int maxc = cgs_max_cols(K->cplx);
  int64_t want = reserve_cols <= 0 ? 128 : reserve_cols;
  K->seg_cols = int(std::max<int64_t>(4, std::min<int64_t>(want, maxc)));
  int rc = [&]() -> int {
    CMB_CUDA(cudaMalloc(&K->v, sizeof(double) * K->ld));
    CMB_CUDA(cudaMalloc(&K->w, sizeof(double) * K->ld));
    CMB_CUDA(cudaMemsetAsync(K->v, 0, sizeof(double) * K->ld, ctx->stream));
    CMB_CUDA(cudaMemsetAsync(K->w, 0, sizeof(double) * K->ld, ctx->stream));
    CMB_CUDA(cudaMalloc(&K->scal, sizeof(double) * 16));
    CMB_CUDA(cudaMemsetAsync(K->scal, 0, sizeof(double) * 16, ctx->stream));
    CMB_CUDA(cudaMalloc(&K->halt, sizeof(int) * 4));
    CMB_CUDA(cudaMemsetAsync(K->halt, 0, sizeof(int) * 4, ctx->stream));
    CMB_CUDA(cudaMalloc(&K->alpha_dev, sizeof(double) * 2 * kMaxSlots));
    CMB_CUDA(cudaMalloc(&K->beta_dev, sizeof(double) * kMaxSlots));
    CMB_CUDA(cudaMalloc(&K->d_idx, sizeof(unsigned long long) * 2));
    CMB_TRY(ensure_stage(K, 4096));
    return CMB_OK;
  }();
  if (rc != CMB_OK) {
    cmb_krylov_destroy(K);
    return rc;
  }
  *out = K;
  return CMB_OK;
}

int cmb_krylov_destroy(cmb_krylov* K) {
  if (!K) return CMB_OK;
  cudaSetDevice(K->ctx->device);
  cudaStreamSynchronize(K->ctx->stream);
  for (auto p : K->segs) dfree(K->ctx, p);
  dfree(K->ctx, K->v);
  dfree(K->ctx, K->w);
  pool_free(K->ctx, K->h1);
  pool_free(K->ctx, K->h2);
  pool_free(K->ctx, K->d_start);
  cudaStreamSynchronize(K->ctx->stream);
  dfree(K->ctx, K->scal);
  dfree(K->ctx, K->halt);
  dfree(K->ctx, K->alpha_dev);
  dfree(K->ctx, K->beta_dev);
  dfree(K->ctx, K->tmp1);
  dfree(K->ctx, K->tmp2);
  dfree(K->ctx, K->tmpz);
  for (int b = 0; b < 2; ++b) {
    dfree(K->ctx, K->xout[b]);
    if (K->ev_x[b]) cudaEventDestroy(K->ev_x[b]);
    if (K->ev_c[b]) cudaEventDestroy(K->ev_c[b]);
  }
  dfree(K->ctx, K->d_idx);
  hfree(K->ctx, K->h_stage);
  delete K;
  return CMB_OK;
}

int cmb_krylov_clear(cmb_krylov* K) {
  CMB_REQUIRE(K, "null argument");
  CMB_CUDA(cudaSetDevice(K->ctx->device));
  K->nk = 0;
  K->started = false;
  K->nrm2_seq = 0;
  K->w_pushed = false;
  K->residue = 0.0;
  K->bytes = 0.0;
  CMB_CUDA(cudaMemsetAsync(K->halt, 0, sizeof(int) * 4, K->ctx->stream));
  return CMB_OK;
}

int cmb_krylov_set_deflation(cmb_krylov* K, int64_t nvec, const void* vecs, int64_t ld) {
  CMB_REQUIRE(K && nvec >= 0 && (nvec == 0 || (vecs && ld >= K->n_local)), "bad deflation vectors");
  CMB_REQUIRE(K->nk == 0, "deflation vectors can only change while the basis is empty");
  CMB_REQUIRE(nvec < kMaxSlots, "too many deflation vectors");
  CMB_CUDA(cudaSetDevice(K->ctx->device));
  K->ndefl = int(nvec);
  if (nvec == 0) return CMB_OK;
  CMB_TRY(ensure_cols(K, int(nvec) + 1));
  const char* src = static_cast<const char*>(vecs);
  for (int64_t j = 0; j < nvec; ++j)
    CMB_CUDA(cudaMemcpyAsync(K->col(int(j)), src + size_t(j) * ld * K->es * sizeof(double),
                             sizeof(double) * K->nd_local, cudaMemcpyHostToDevice, K->ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(K->ctx->stream));
  return CMB_OK;
}

// init: host vector, or nullptr = the copy of the previous start vector kept in K->d_start
static int krylov_start(cmb_krylov* K, const void* init, double threshold, int* status) {
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  K->nk = 0;
  K->started = false;
  K->nrm2_seq = 0;
  K->w_pushed = false;
  K->residue = 0.0;
  CMB_CUDA(cudaMemsetAsync(K->halt, 0, sizeof(int) * 4, ctx->stream));
  if (init) {
    if (!K->d_start) CMB_TRY(pool_alloc(ctx, &K->d_start, sizeof(double) * std::max<int64_t>(K->nd_local, 2)));
    CMB_CUDA(cudaMemcpyAsync(K->d_start, init, sizeof(double) * K->nd_local, cudaMemcpyHostToDevice, ctx->stream));
  }
  CMB_CUDA(cudaMemcpyAsync(K->w, K->d_start, sizeof(double) * K->nd_local, cudaMemcpyDeviceToDevice, ctx->stream));
  if (K->ndefl > 0) {
    // lanczos.hpp:312-314: project the deflation vectors out of the start vector
    std::vector<Chunk> chunks;
    contiguous_chunks(K, 0, K->ndefl, chunks);
    CMB_TRY(gram_schmidt2(K, chunks, K->w, K->w, K->scal));
  } else {
    CMB_TRY(vec_dot(ctx, K->cplx, K->w, K->w, K->ld, K->scal + 2, K->halt));
    CMB_CUDA(cudaMemcpyAsync(K->scal, K->scal + 2, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CMB_TRY(allreduce_sum_f64(ctx, K->scal, 1));
  }
  CMB_CUDA(cudaMemcpyAsync(K->h_stage, K->scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  const double nrm = sqrt(K->h_stage[0]);
  if (!(nrm >= threshold)) {  // lanczos.hpp:316-318 (`nrm < threshold_` -> no vector); NaN also fails
    *status = CMB_STEP_NOSTART;
    return CMB_OK;
  }
  K->started = true;
  *status = CMB_STEP_OK;
  return CMB_OK;
}

int cmb_krylov_start(cmb_krylov* K, const void* init, double threshold, int* status) {
  CMB_REQUIRE(K && init && status, "null argument");
  return krylov_start(K, init, threshold, status);
}

int cmb_krylov_restart(cmb_krylov* K, double threshold, int* status) {
  CMB_REQUIRE(K && status, "null argument");
  CMB_REQUIRE(K->d_start, "cmb_krylov_restart needs an earlier cmb_krylov_start on this state");
  return krylov_start(K, nullptr, threshold, status);
}

int64_t cmb_krylov_ncols(const cmb_krylov* K) { return K ? K->nk : 0; }
int64_t cmb_krylov_rows(const cmb_krylov* K) { return K ? K->n_local : 0; }
double cmb_krylov_bytes(const cmb_krylov* K) { return K ? K->bytes : 0.0; }

int cmb_krylov_get_col(cmb_krylov* K, int64_t j, void* out) {
  CMB_REQUIRE(K && out && j >= 0 && j < K->nk, "column index out of range");
  CMB_CUDA(cudaSetDevice(K->ctx->device));
  CMB_CUDA(cudaMemcpyAsync(out, K->col(K->ndefl + int(j)), sizeof(double) * K->nd_local, cudaMemcpyDeviceToHost,
                           K->ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(K->ctx->stream));
  return CMB_OK;
}

int cmb_lanczos_run(cmb_krylov* K, cmb_op* op, double shift, int64_t interval, double threshold, int64_t nsteps,
                    double* alpha, double* beta, int64_t* steps_done, int* status) {
  CMB_TRY(check_pair(K, op));
  CMB_REQUIRE(alpha && beta && steps_done && status && nsteps >= 0, "bad argument");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  if (ctx->dead) return check_peer_wait(ctx);
  *steps_done = 0;
  *status = CMB_STEP_OK;
  if (nsteps == 0) return CMB_OK;
  const int nk0 = K->nk;
  CMB_REQUIRE(K->ndefl + nk0 + nsteps + 1 < kMaxSlots, "too many Krylov steps");
  int enq = 0;  // steps enqueued
  for (int64_t s = 0; s < nsteps; ++s) {
    StepScalars sc;
    sc.nrm2 = K->scal;
    sc.halt = K->halt;
    if (K->nk + enq == 0) {
      // first call (lanczos.hpp:378-398): u0 = start/||start|| ; v = (A+shift) u0 ; alpha0
      CMB_REQUIRE(K->started, "cmb_krylov_start must succeed before the first step");
      CMB_TRY(ensure_cols(K, K->ndefl + 1));
      sc.threshold = -1.0;
      sc.beta_slot = K->scal + 3;
      sc.alpha_slot = K->alpha_dev;
      CMB_TRY(apply_w(K, op, K->col(K->ndefl), shift, 0.0, sc));
      if (interval != 1) CMB_TRY(allreduce_sum_f64(ctx, K->alpha_dev, 1));
      K->bytes += op->bytes + 3.0 * double(K->n_local) * (K->cplx ? 16.0 : 8.0);
      K->nk = 1;  // the first call cannot break down on the device side
      continue;
    }
    const int k = K->nk + enq - 1;
    CMB_TRY(ensure_cols(K, K->ndefl + k + 2));
    const int nk_save = K->nk;
    K->nk = k + 1;  // enqueue_lanczos_orth works on the state "k+1 vectors"
    int rc = CMB_OK;
    if (K->resume_norm_ready) {
      // resumed after the Pythagorean guard: w is orthogonalised already, scal[0] holds the explicit norm
      K->resume_norm_ready = false;
      K->nrm2_seq = 0;
    } else {
      rc = enqueue_lanczos_orth(K, interval, enq, op);
    }
    const int c = orth_cols_count(K, interval);
    K->nk = nk_save;
    CMB_TRY(rc);
    sc.threshold = threshold;
    if (K->nrm2_seq) {
      sc.nrm2_pull = mail_pull_of(ctx, K->nrm2_seq, K->scal);  // ||w||^2 sits in the mailbox
      sc.nrm2_pull.halt = K->halt;
    }
    sc.beta_slot = K->beta_dev + k;
    sc.alpha_slot = K->alpha_dev + size_t(k + 1) * 2;
    CMB_TRY(apply_w(K, op, K->col(K->ndefl + k + 1), shift, 0.0, sc));
    // with full reorthogonalisation alpha is only read by the host: one allreduce for the whole chain (below)
    if (interval != 1) CMB_TRY(allreduce_sum_f64(ctx, K->alpha_dev + size_t(k + 1) * 2, 1));
    add_step_bytes(K, op, c);
    ++enq;
  }
  // one synchronisation for the whole chain
  const bool did_first = (nk0 == 0);
  const int a0 = did_first ? 0 : nk0;  // first alpha slot produced
  const int na = (did_first ? 1 : 0) + enq;
  const int b0 = nk0 == 0 ? 0 : nk0 - 1;
  const int nb = enq;
  CMB_TRY(ensure_stage(K, size_t(2 * na + nb + 8)));
  double* hs = K->h_stage;
  if (interval == 1 && na) CMB_TRY(allreduce_sum_f64(ctx, K->alpha_dev + size_t(a0) * 2, size_t(2) * na));
  if (na) CMB_CUDA(cudaMemcpyAsync(hs, K->alpha_dev + size_t(a0) * 2, sizeof(double) * 2 * na, cudaMemcpyDeviceToHost, ctx->stream));
  if (nb) CMB_CUDA(cudaMemcpyAsync(hs + 2 * na, K->beta_dev + b0, sizeof(double) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  CMB_CUDA(cudaMemcpyAsync(hs + 2 * na + nb, K->halt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  const int halted = *reinterpret_cast<int*>(hs + 2 * na + nb);
  CMB_TRY(check_peer_wait(ctx));
  int ok_steps = enq;
  bool guard_retry = false;
  if (halted) K->w_pushed = false;  // pushes announced after the halt never ran
  if (halted) {
    int flags[3] = {0, 0, 0};
    CMB_TRY(d2h_sync(ctx, flags, K->halt, sizeof(flags)));
    guard_retry = flags[1] != 0;
    CMB_CUDA(cudaMemsetAsync(K->halt, 0, sizeof(int) * 4, ctx->stream));
    if (guard_retry) {
      ok_steps = flags[2];  // steps of this chain completed before the guarded one
    } else {
      ok_steps = 0;
      while (ok_steps < enq && hs[2 * na + ok_steps] > threshold) ++ok_steps;
      *status = CMB_STEP_BREAKDOWN;
    }
  }
  int ai = 0;
  if (did_first) alpha[ai++] = hs[0];
  for (int s = 0; s < ok_steps; ++s) alpha[ai++] = hs[2 * ((did_first ? 1 : 0) + s)];
  const int nb_out = (halted && !guard_retry) ? ok_steps + 1 : ok_steps;  // the breaking beta is kept (lanczos.hpp:433-436)
  for (int s = 0; s < nb_out && s < enq; ++s) beta[s] = hs[2 * na + s];
  K->nk = (did_first ? 1 : nk0) + ok_steps;
  *steps_done = (did_first ? 1 : 0) + ok_steps;
  if (guard_retry) {
    // The Pythagorean beta^2 = ||w1||^2 - |h2|^2 of step ok_steps lost its digits to cancellation.  w itself is fine
    // (the third pass has written it): reduce ||w||^2 explicitly and resume the chain with the remaining steps.
    CMB_TRY(vec_dot(ctx, K->cplx, K->w, K->w, K->ld, K->scal + 2, K->halt));
    CMB_CUDA(cudaMemcpyAsync(K->scal, K->scal + 2, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CMB_TRY(allreduce_sum_f64(ctx, K->scal, 1));
    K->resume_norm_ready = true;
    const int64_t done_here = *steps_done;
    int64_t more = 0;
    int rc2 = cmb_lanczos_run(K, op, shift, interval, threshold, nsteps - done_here, alpha + ai, beta + nb_out, &more, status);
    K->resume_norm_ready = false;
    CMB_TRY(rc2);
    *steps_done = done_here + more;
  }
  return CMB_OK;
}

int cmb_lanczos_step(cmb_krylov* K, cmb_op* op, double shift, int64_t interval, double threshold, double* alpha,
                     double* beta, int* status) {
  CMB_REQUIRE(alpha && beta && status, "null argument");
  int64_t done = 0;
  double a[2] = {0, 0}, b[2] = {0, 0};
  CMB_TRY(cmb_lanczos_run(K, op, shift, interval, threshold, 1, a, b, &done, status));
  *alpha = a[0];
  *beta = b[0];
  return CMB_OK;
}

static int ensure_tmp(cmb_krylov* K, double** p, size_t doubles);

// Diagnostic: time `reps` launches of one Gram-Schmidt pass (mode 0/1/2) over the first `ncols` basis columns
// with CUDA events (columns are allocated on demand; their contents do not matter for timing).
int cmb_debug_cgs_pass(cmb_krylov* K, int mode, int ncols, int reps, double* ms_per_launch) {
  CMB_REQUIRE(K && ms_per_launch && mode >= 0 && mode <= 2 && ncols >= 1 && reps >= 1, "bad argument");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  CMB_TRY(ensure_cols(K, ncols));
  std::vector<Chunk> chunks;
  contiguous_chunks(K, 0, ncols, chunks);
  CMB_REQUIRE(chunks.size() == 1, "diagnostic pass needs a single chunk (ncols <= segment size)");
  CMB_CUDA(cudaMemsetAsync(K->h1, 0, sizeof(double) * 2 * ncols, ctx->stream));
  CgsPass p;
  p.ld = K->ld;
  p.halt = K->halt;
  p.V = chunks[0].V;
  p.ncols = ncols;
  p.col_stride = chunks[0].col_stride;
  p.x = K->v;
  p.y = (mode == 0) ? nullptr : K->w;
  p.hin = (mode == 0) ? nullptr : K->h1;
  p.hout = (mode == 2) ? K->scal + 1 : K->h2;
  CMB_TRY(cgs_pass(ctx, K->cplx, mode, p));  // warm-up
  cudaEvent_t a, b;
  CMB_CUDA(cudaEventCreate(&a));
  CMB_CUDA(cudaEventCreate(&b));
  CMB_CUDA(cudaEventRecord(a, ctx->stream));
  for (int r = 0; r < reps; ++r) CMB_TRY(cgs_pass(ctx, K->cplx, mode, p));
  CMB_CUDA(cudaEventRecord(b, ctx->stream));
  CMB_CUDA(cudaEventSynchronize(b));
  float ms = 0.f;
  CMB_CUDA(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *ms_per_launch = double(ms) / reps;
  return CMB_OK;
}

int cmb_lanczos_residual_norm(cmb_krylov* K, double* out) {
  CMB_REQUIRE(K && out, "null argument");
  CMB_REQUIRE(K->nk >= 1, "no Lanczos vector yet");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  CMB_TRY(ensure_tmp(K, &K->tmp1, K->ld));
  const int k = K->nk - 1;
  const int first = (k > 0) ? k - 1 : k;
  std::vector<Chunk> chunks;
  contiguous_chunks(K, K->ndefl + first, K->ndefl + k + 1, chunks);
  CMB_CUDA(cudaMemsetAsync(K->h2, 0, sizeof(double) * 4, ctx->stream));
  int slot = 0;
  if (k > 0) {
    CMB_CUDA(cudaMemcpyAsync(K->h2 + slot * K->es, K->beta_dev + (k - 1), sizeof(double), cudaMemcpyDeviceToDevice,
                             ctx->stream));
    ++slot;
  }
  CMB_CUDA(cudaMemcpyAsync(K->h2 + slot * K->es, K->alpha_dev + size_t(k) * 2, sizeof(double), cudaMemcpyDeviceToDevice,
                           ctx->stream));
  CMB_TRY(subtract_cols(K, chunks, K->h2, K->v, K->tmp1, K->scal + 1, "recurrence"));
  CMB_CUDA(cudaMemcpyAsync(K->h_stage, K->scal + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = sqrt(K->h_stage[0]);
  return CMB_OK;
}

// Thick restart (Wu & Simon 2000), additive.  State on entry: Krylov vectors u_0..u_m (nk = m+1), v = (A+shift) u_m,
// alpha_m known.  The caller diagonalised the projected matrix of the first m vectors and passes the m x nkeep
// coefficients of the Ritz vectors to keep.  The basis becomes [V_m coef, u_m] (nkeep + 1 vectors); v and alpha_m stay,
// so the next Lanczos step (full reorthogonalisation) continues from u_m against the compressed basis.
int cmb_lanczos_thick_restart(cmb_krylov* K, const double* coef, int64_t ldc, int64_t m, int64_t nkeep) {
  CMB_REQUIRE(K && coef, "null argument");
  CMB_REQUIRE(m >= 1 && m == K->nk - 1 && nkeep >= 1 && nkeep <= m && ldc >= m, "bad shape for a thick restart");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  const int nk = K->nk;
  CMB_TRY(ensure_cols(K, K->ndefl + nk + int(nkeep)));  // scratch columns behind the basis
  std::vector<Chunk> chunks;
  contiguous_chunks(K, K->ndefl, K->ndefl + int(m), chunks);
  const int es = K->es;
  CMB_TRY(ensure_stage(K, size_t(nkeep) * size_t(m) * es + 8));
  for (int64_t i = 0; i < nkeep; ++i) {
    double* hs = K->h_stage + size_t(i) * size_t(m) * es;
    for (int64_t r = 0; r < m; ++r) {
      hs[r * es] = -coef[size_t(i) * ldc + r];
      if (es == 2) hs[r * es + 1] = 0.0;
    }
    CMB_CUDA(cudaMemcpyAsync(K->h1, hs, sizeof(double) * m * es, cudaMemcpyHostToDevice, ctx->stream));
    CMB_TRY(subtract_cols(K, chunks, K->h1, nullptr, K->col(K->ndefl + nk + int(i)), K->scal + 1, "ritz_assemble"));
    K->bytes += (double(m) + 1.0) * double(K->n_local) * (K->cplx ? 16.0 : 8.0);
  }
  const size_t colbytes = sizeof(double) * size_t(K->ld);
  for (int64_t i = 0; i < nkeep; ++i)
    CMB_CUDA(cudaMemcpyAsync(K->col(K->ndefl + int(i)), K->col(K->ndefl + nk + int(i)), colbytes, cudaMemcpyDeviceToDevice,
                             ctx->stream));
  if (nkeep != m) {
    CMB_CUDA(cudaMemcpyAsync(K->col(K->ndefl + int(nkeep)), K->col(K->ndefl + int(m)), colbytes, cudaMemcpyDeviceToDevice,
                             ctx->stream));
    CMB_CUDA(cudaMemcpyAsync(K->alpha_dev + size_t(nkeep) * 2, K->alpha_dev + size_t(m) * 2, sizeof(double) * 2,
                             cudaMemcpyDeviceToDevice, ctx->stream));
  }
  K->nk = int(nkeep) + 1;
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));  // the staging slices are reused by the next call
  return CMB_OK;
}

// Thick restart of the Arnoldi iteration (Krylov-Schur style; Stewart, SIAM J. Matrix Anal. Appl. 23 (2001)).  State on
// entry: Arnoldi vectors q_0..q_{m-1} (nk = m), w = the unnormalised residual vector of the last step with ||w|| =
// residue.  coef is the m x nkeep matrix (column-major, leading dimension ldc, elements of the basis dtype) of an
// orthonormal basis Z of the subspace to keep; the basis becomes Q Z (nkeep vectors), w and its norm stay, so the next
// cmb_arnoldi_run step turns w into vector number nkeep and orthogonalises A q against the compressed basis.  The caller
// keeps the projected matrix: Z^H H Z in the leading block and residue * Z(m-1, :) as row nkeep.
int cmb_arnoldi_thick_restart(cmb_krylov* K, const void* coef, int64_t ldc, int64_t m, int64_t nkeep) {
  CMB_REQUIRE(K && coef, "null argument");
  CMB_REQUIRE(m >= 1 && m == K->nk && nkeep >= 1 && nkeep <= m && ldc >= m, "bad shape for a thick restart");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  if (ctx->dead) return check_peer_wait(ctx);
  CMB_TRY(ensure_cols(K, K->ndefl + int(m) + int(nkeep)));  // scratch columns behind the basis
  std::vector<Chunk> chunks;
  contiguous_chunks(K, K->ndefl, K->ndefl + int(m), chunks);
  const int es = K->es;
  CMB_TRY(ensure_stage(K, size_t(nkeep) * size_t(m) * es + 8));
  const double* cf = static_cast<const double*>(coef);
  // the device scalars of the pending residual vector must survive the assembly passes (they use scal[1] only)
  for (int64_t i = 0; i < nkeep; ++i) {
    double* hs = K->h_stage + size_t(i) * size_t(m) * es;
    for (int64_t r = 0; r < m * es; ++r) hs[r] = -cf[size_t(i) * ldc * es + r];  // y = 0 - V (-z_i)
    CMB_CUDA(cudaMemcpyAsync(K->h1, hs, sizeof(double) * m * es, cudaMemcpyHostToDevice, ctx->stream));
    CMB_TRY(subtract_cols(K, chunks, K->h1, nullptr, K->col(K->ndefl + int(m) + int(i)), K->scal + 1, "ritz_assemble"));
    K->bytes += (double(m) + 1.0) * double(K->n_local) * (K->cplx ? 16.0 : 8.0);
  }
  const size_t colbytes = sizeof(double) * size_t(K->ld);
  for (int64_t i = 0; i < nkeep; ++i)
    CMB_CUDA(cudaMemcpyAsync(K->col(K->ndefl + int(i)), K->col(K->ndefl + int(m) + int(i)), colbytes, cudaMemcpyDeviceToDevice,
                             ctx->stream));
  K->nk = int(nkeep);
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));  // the staging slices are reused by the next call
  return CMB_OK;
}

int cmb_arnoldi_run(cmb_krylov* K, cmb_op* op, const void* shift, double threshold, int64_t nsteps, void* hcols,
                    int64_t ldh, double* residues, int64_t* steps_done, int* status) {
  CMB_TRY(check_pair(K, op));
  CMB_REQUIRE(hcols && residues && steps_done && status && nsteps >= 0, "bad argument");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  if (ctx->dead) return check_peer_wait(ctx);
  double shr = 0.0, shi = 0.0;
  if (shift) {
    shr = static_cast<const double*>(shift)[0];
    if (K->cplx) shi = static_cast<const double*>(shift)[1];
  }
  *status = CMB_STEP_OK;
  *steps_done = 0;
  if (nsteps == 0) return CMB_OK;
  if (K->nk == 0) {
    CMB_REQUIRE(K->started, "cmb_krylov_start must succeed before the first step");
  } else {
    // arnoldiStepIsUtmost (arnoldi.hpp:277-288)
    if (K->nk >= K->n_global) {
      *status = CMB_STEP_FULL;
      return CMB_OK;
    }
    if (K->residue <= threshold) {
      *status = CMB_STEP_BREAKDOWN;
      return CMB_OK;
    }
  }
  const int k0 = K->nk;
  if (nsteps > K->n_global - k0) nsteps = K->n_global - k0;  // the basis cannot exceed the dimension
  CMB_REQUIRE(K->ndefl + k0 + nsteps + 2 < kMaxSlots, "too many Krylov steps");
  CMB_REQUIRE(ldh >= k0 + nsteps, "ldh too small for the Hessenberg columns");
  const int es = K->es;
  const int cmax = K->ndefl + k0 + int(nsteps);
  // per-step history of h1, h2 and ||w||^2 so that the whole chain needs one host synchronisation
  const size_t hstride = size_t(cmax) * es;
  double* hist = nullptr;
  CMB_TRY(pool_alloc(ctx, &hist, sizeof(double) * (2 * hstride + 1) * size_t(nsteps)));
  int rc = CMB_OK;
  for (int64_t s = 0; s < nsteps && rc == CMB_OK; ++s) {
    const int k = k0 + int(s);  // index of the vector created now
    rc = ensure_cols(K, K->ndefl + k + 1);
    if (rc != CMB_OK) break;
    StepScalars sc;
    sc.nrm2 = K->scal;
    sc.halt = K->halt;
    // the host has tested the residue of the state it knows; later steps of the chain test it on the device
    sc.threshold = (s == 0) ? -1.0 : threshold;
    sc.beta_slot = K->scal + 3;
    sc.alpha_slot = K->scal + 4;
    // q_k = w / residue ; v = (A + shift) q_k        (arnoldi.hpp:361-372)
    rc = apply_w(K, op, K->col(K->ndefl + k), shr, shi, sc);
    if (rc != CMB_OK) break;
    // CGS2 of v against deflation vectors and q_0..q_k ; h(:,k) = h1 + h2 ; residue = ||w||   (:373-385)
    std::vector<Chunk> chunks;
    const int c = K->ndefl + k + 1;
    contiguous_chunks(K, 0, c, chunks);
    if (ctx->mail_ok && chunks.size() == 1 && K->ndefl == 0) {
      // row-partitioned fast path: the coefficient reductions go through the peer-memory mailboxes (reduced h1, h2
      // are written back to K->h1 / K->h2, ||w||^2 to K->scal[0]) — no collective kernel inside the chain
      unsigned long long norm_seq = 0;  // stays 0 without deflation vectors: the norm is in K->scal[0]
      rc = gram_schmidt2_mailed(K, chunks[0], K->v, K->w, &norm_seq, int(s), op);
    } else {
      rc = gram_schmidt2(K, chunks, K->v, K->w, K->scal, op);
    }
    if (rc != CMB_OK) break;
    double* slot = hist + size_t(s) * (2 * hstride + 1);
    cudaMemcpyAsync(slot, K->h1, sizeof(double) * c * es, cudaMemcpyDeviceToDevice, ctx->stream);
    cudaMemcpyAsync(slot + hstride, K->h2, sizeof(double) * c * es, cudaMemcpyDeviceToDevice, ctx->stream);
    cudaMemcpyAsync(slot + 2 * hstride, K->scal, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
    add_step_bytes(K, op, c);
  }
  if (rc == CMB_OK) rc = ensure_stage(K, (2 * hstride + 1) * size_t(nsteps) + 8);
  int halted = 0;
  if (rc == CMB_OK) {
    double* hs = K->h_stage;
    cudaError_t e = cudaMemcpyAsync(hs, hist, sizeof(double) * (2 * hstride + 1) * size_t(nsteps), cudaMemcpyDeviceToHost,
                                    ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(hs + (2 * hstride + 1) * size_t(nsteps), K->halt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("cmb_arnoldi_run: %s", cudaGetErrorString(e));
      rc = CMB_ERR_CUDA;
    } else {
      halted = *reinterpret_cast<int*>(hs + (2 * hstride + 1) * size_t(nsteps));
    }
  }
  pool_free(ctx, hist);
  CMB_TRY(rc);
  CMB_TRY(check_peer_wait(ctx));
  int guard_step = -1;
  if (halted) K->w_pushed = false;  // pushes announced after the halt never ran
  if (halted) {
    int flags[3] = {0, 0, 0};
    CMB_TRY(d2h_sync(ctx, flags, K->halt, sizeof(flags)));
    if (flags[1]) guard_step = flags[2];
  }
  // steps are valid until a residue <= threshold appears (that step is still valid; the next one was refused)
  double* hs = K->h_stage;
  int64_t done = 0;
  for (int64_t s = 0; s < nsteps; ++s) {
    const double* slot = hs + size_t(s) * (2 * hstride + 1);
    const int k = k0 + int(s);
    double* hout = static_cast<double*>(hcols) + size_t(s) * ldh * es;
    for (int i = 0; i < (k + 1) * es; ++i) hout[i] = slot[K->ndefl * es + i] + slot[hstride + K->ndefl * es + i];
    residues[s] = sqrt(slot[2 * hstride]);
    ++done;
    if (int(s) == guard_step) break;  // the Pythagorean residue of this step is unreliable: replaced below
    if (residues[s] <= threshold) break;
  }
  if (halted) CMB_CUDA(cudaMemsetAsync(K->halt, 0, sizeof(int) * 4, ctx->stream));
  if (guard_step >= 0) {
    // cancellation in ||w1||^2 - |h2|^2: w is orthogonalised, reduce its norm explicitly and go on with the chain
    CMB_TRY(vec_dot(ctx, K->cplx, K->w, K->w, K->ld, K->scal + 2, K->halt));
    CMB_CUDA(cudaMemcpyAsync(K->scal, K->scal + 2, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CMB_TRY(allreduce_sum_f64(ctx, K->scal, 1));
    CMB_CUDA(cudaMemcpyAsync(K->h_stage, K->scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    residues[done - 1] = sqrt(K->h_stage[0]);
    K->residue = residues[done - 1];
    K->nk = k0 + int(done);
    *steps_done = done;
    if (done < nsteps) {
      int64_t more = 0;
      CMB_TRY(cmb_arnoldi_run(K, op, shift, threshold, nsteps - done, static_cast<double*>(hcols) + size_t(done) * ldh * es, ldh,
                              residues + done, &more, status));
      *steps_done = done + more;
    }
    return CMB_OK;
  }
  if (halted) {
    *status = CMB_STEP_BREAKDOWN;
    if (done < nsteps) {
      // the chain stopped on the device: w and ||w||^2 still describe the last valid step
      CMB_CUDA(cudaMemcpyAsync(K->scal, hs + size_t(done - 1) * (2 * hstride + 1) + 2 * hstride, sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
      CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
  }
  K->residue = residues[done - 1];
  K->nk = k0 + int(done);
  *steps_done = done;
  return CMB_OK;
}

int cmb_arnoldi_step(cmb_krylov* K, cmb_op* op, const void* shift, double threshold, void* hcol, double* residue,
                     int* status) {
  CMB_REQUIRE(K && hcol && residue && status, "null argument");
  int64_t done = 0;
  double res = K->residue;
  CMB_TRY(cmb_arnoldi_run(K, op, shift, threshold, 1, hcol, K->nk + 2, &res, &done, status));
  *residue = done ? res : K->residue;
  return CMB_OK;
}

}  // extern "C"

static int ensure_tmp(cmb_krylov* K, double** p, size_t doubles) {
  if (*p) return CMB_OK;
  if (cudaMalloc(p, sizeof(double) * doubles) != cudaSuccess) {
    cudaGetLastError();
    set_error("out of device memory for a Ritz-vector work buffer");
    return CMB_ERR_NOMEM;
  }
  CMB_CUDA(cudaMemsetAsync(*p, 0, sizeof(double) * doubles, K->ctx->stream));
  return CMB_OK;
}

__global__ void interleave_kernel(const double* __restrict__ re, const double* __restrict__ im, double* __restrict__ z,
                                  long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    reinterpret_cast<double2*>(z)[i] = make_double2(re[i], im[i]);
}
__global__ void add2_kernel(const double* a, const double* b, double* out) { out[0] = a[0] + b[0]; }

extern "C" {

// g = V^H x over the Krylov vectors (x: local host slab) — the projection LanczosExponentialSolver needs
// (lanczos.hpp:1047: inner = y_n^H in, with y_n = V S(:,n)).
int cmb_krylov_project(cmb_krylov* K, const void* x_host, void* g_host) {
  CMB_REQUIRE(K && x_host && g_host, "null argument");
  CMB_REQUIRE(K->nk >= 1, "empty basis");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  CMB_TRY(ensure_tmp(K, &K->tmp1, K->ld));
  CMB_CUDA(cudaMemcpyAsync(K->tmp1, x_host, sizeof(double) * K->nd_local, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<Chunk> chunks;
  contiguous_chunks(K, K->ndefl, K->ndefl + K->nk, chunks);
  CgsPass p;
  p.ld = K->ld;
  p.halt = K->halt;
  p.x = K->tmp1;
  p.family = "project";
  int off = 0;
  for (auto& c : chunks) {
    p.V = c.V;
    p.ncols = c.ncols;
    p.col_stride = c.col_stride;
    p.hout = K->h1 + off * K->es;
    CMB_TRY(cgs_pass(ctx, K->cplx, CGS_DOT, p));
    off += c.ncols;
  }
  CMB_TRY(allreduce_sum_f64(ctx, K->h1, size_t(K->nk) * K->es));
  CMB_CUDA(cudaMemcpyAsync(g_host, K->h1, sizeof(double) * K->nk * K->es, cudaMemcpyDeviceToHost, ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CMB_OK;
}

// out = sum_m coef_m * u_m over the first ncoef Krylov vectors (no normalisation, no phase): the tall-skinny
// GEMV that turns a small vector of the Krylov space into a length-n vector.
int cmb_krylov_combine(cmb_krylov* K, const void* coef, int64_t ncoef, void* out_host) {
  CMB_REQUIRE(K && coef && out_host, "null argument");
  CMB_REQUIRE(ncoef >= 1 && ncoef <= K->nk, "bad coefficient count");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  CMB_TRY(ensure_tmp(K, &K->tmp1, K->ld));
  CMB_TRY(ensure_stage(K, size_t(2 * ncoef + 8)));
  const double* cf = static_cast<const double*>(coef);
  for (int64_t m = 0; m < ncoef * K->es; ++m) K->h_stage[m] = -cf[m];  // y = 0 - V (-coef)
  CMB_CUDA(cudaMemcpyAsync(K->h1, K->h_stage, sizeof(double) * ncoef * K->es, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<Chunk> chunks;
  contiguous_chunks(K, K->ndefl, K->ndefl + int(ncoef), chunks);
  CMB_TRY(subtract_cols(K, chunks, K->h1, nullptr, K->tmp1, K->scal + 1, "combine"));
  CMB_CUDA(cudaMemcpyAsync(out_host, K->tmp1, sizeof(double) * K->nd_local, cudaMemcpyDeviceToHost, ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CMB_OK;
}

int cmb_krylov_ritz_vectors(cmb_krylov* K, cmb_dtype coef_dtype, const void* coef, int64_t ldc, int64_t ncoef,
                            int64_t nev, void* x_host, int64_t ldx) {
  CMB_REQUIRE(K && (nev == 0 || (coef && x_host)), "null argument");
  CMB_REQUIRE(ncoef >= 0 && ncoef <= K->nk && ldc >= ncoef && ldx >= K->n_local && nev >= 0, "bad shape");
  CMB_REQUIRE(coef_dtype == CMB_F64 || coef_dtype == CMB_C64, "bad coefficient dtype");
  const bool ccplx = coef_dtype == CMB_C64;
  CMB_REQUIRE(!(K->cplx && !ccplx), "a complex basis needs complex coefficients");
  cmb_ctx* ctx = K->ctx;
  CMB_CUDA(cudaSetDevice(ctx->device));
  if (nev == 0) return CMB_OK;
  const int ces = ccplx ? 2 : 1;
  const bool widen = ccplx && !K->cplx;  // real basis, complex coefficients (ArnoldiEigenSolver<double>)
  if (widen) {
    CMB_TRY(ensure_tmp(K, &K->tmp1, K->ld));
    CMB_TRY(ensure_tmp(K, &K->tmp2, K->ld));
  }
  const size_t out_need = size_t(widen ? 2 * K->ld : K->ld);
  if (K->xout_doubles < out_need) {
    CMB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < 2; ++b) {
      dfree(ctx, K->xout[b]);
      K->xout[b] = nullptr;
    }
    K->xout_doubles = 0;
  }
  for (int b = 0; b < 2; ++b) {
    CMB_TRY(ensure_tmp(K, &K->xout[b], out_need));
    if (!K->ev_x[b]) CMB_CUDA(cudaEventCreateWithFlags(&K->ev_x[b], cudaEventDisableTiming));
    if (!K->ev_c[b]) CMB_CUDA(cudaEventCreateWithFlags(&K->ev_c[b], cudaEventDisableTiming));
  }
  K->xout_doubles = std::max(K->xout_doubles, out_need);
  std::vector<Chunk> chunks;
  contiguous_chunks(K, K->ndefl, K->ndefl + int(ncoef), chunks);
  CMB_TRY(ensure_stage(K, size_t(nev) * size_t(4 * ncoef + 8)));  // one staging slice per vector: no host sync in the loop
  const double* cf = static_cast<const double*>(coef);
  const size_t out_es = ccplx ? 2 : 1;
  const size_t out_ld = out_need;
  for (int64_t e = 0; e < nev; ++e) {
    double* hs = K->h_stage + size_t(e) * size_t(4 * ncoef + 8);
    const double* ce = cf + size_t(e) * ldc * ces;
    const int b = int(e & 1);
    double* xdev = K->xout[b];
    if (e >= 2) CMB_CUDA(cudaStreamWaitEvent(ctx->stream, K->ev_c[b], 0));  // its previous content is on the host
    if (ncoef == 0) {
      CMB_CUDA(cudaMemsetAsync(xdev, 0, sizeof(double) * out_ld, ctx->stream));
      CMB_CUDA(cudaMemsetAsync(K->scal + 1, 0, sizeof(double), ctx->stream));
    } else if (!widen) {
      // x = 0 - V (-coef)
      for (int64_t m = 0; m < ncoef * ces; ++m) hs[m] = -ce[m];
      CMB_CUDA(cudaMemcpyAsync(K->h1, hs, sizeof(double) * ncoef * ces, cudaMemcpyHostToDevice, ctx->stream));
      CMB_TRY(subtract_cols(K, chunks, K->h1, nullptr, xdev, K->scal + 1, "ritz_assemble"));
    } else {
      for (int64_t m = 0; m < ncoef; ++m) {
        hs[m] = -ce[2 * m];
        hs[ncoef + m] = -ce[2 * m + 1];
      }
      CMB_CUDA(cudaMemcpyAsync(K->h1, hs, sizeof(double) * 2 * ncoef, cudaMemcpyHostToDevice, ctx->stream));
      CMB_TRY(subtract_cols(K, chunks, K->h1, nullptr, K->tmp1, K->scal + 2, "ritz_assemble"));
      CMB_TRY(subtract_cols(K, chunks, K->h1 + ncoef, nullptr, K->tmp2, K->scal + 3, "ritz_assemble"));
      {
        LaunchScope ls(ctx, "vec_scale");
        const int grid = int(std::max<long long>(1, std::min<long long>((K->ld + 255) / 256, (long long)ctx->num_sms * 8)));
        interleave_kernel<<<grid, 256, 0, ctx->stream>>>(K->tmp1, K->tmp2, xdev, K->ld);
        add2_kernel<<<1, 1, 0, ctx->stream>>>(K->scal + 2, K->scal + 3, K->scal + 1);
        ctx->launches++;
      }
      CMB_CUDA(cudaGetLastError());
    }
    // phase of the first non-zero element (lanczos.hpp:806-813) and normalisation (:816)
    // global index of the first non-zero element (min over ranks); its owner contributes the value, which is
    // copied aside because the scaling kernel overwrites it while other CTAs still read it
    CMB_TRY(vec_first_nonzero(ctx, ccplx, xdev, K->n_local, K->row_begin, K->d_idx));
    CMB_TRY(allreduce_min_u64(ctx, K->d_idx, 1));
    CMB_TRY(vec_pick_element(ctx, xdev, K->d_idx, K->row_begin, K->n_local, int(out_es), K->scal + 6));
    CMB_TRY(allreduce_sum_f64(ctx, K->scal + 6, out_es));
    const double* phase_src = K->scal + 6;  // (0,0) when the vector is identically zero: no phase change
    CMB_TRY(vec_scale_phase(ctx, ccplx, xdev, K->scal + 1, phase_src, out_ld));
    // device -> host on the copy stream: overlaps the assembly of the next vector
    CMB_CUDA(cudaEventRecord(K->ev_x[b], ctx->stream));
    CMB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, K->ev_x[b], 0));
    CMB_CUDA(cudaMemcpyAsync(static_cast<char*>(x_host) + size_t(e) * ldx * out_es * sizeof(double), xdev,
                             sizeof(double) * K->n_local * out_es, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CMB_CUDA(cudaEventRecord(K->ev_c[b], ctx->copy_stream));
  }
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  return CMB_OK;
}

}  // extern "C"
