// heisenberg.cu — matrix-free spin-1/2 Heisenberg chain  H = J sum_i [SzSz + (S+S- + S-S+)/2]_{i,i+1}
// (BASELINE cfg 5; the reference has no such operator).  Basis = bit strings, bit i = spin i;
// y_s = J/4 (#aligned - #anti-aligned bonds) x_s + J/2 sum_{anti-aligned bonds (i,j)} x_{s ^ (1<<i | 1<<j)}.
//
// Row partition over P = 2^p ranks: the top p bits of the state index are the rank, the low Ll = L - p bits
// the local index.  Bonds fall into four classes:
//   local      i, j < Ll                      gathers from the local slab
//   straddle   (Ll-1, Ll)                     partner rank^1, local index s ^ (1<<(Ll-1)): the partner's half slab
//                                             whose top local bit equals my rank bit 0 (contiguous)
//   rank-rank  i >= Ll, j = i+1 < L            uniform per rank: if my two rank bits differ, the whole slab of
//                                             rank ^ (3 << (i-Ll)) is needed, otherwise nothing
//   wrap (PBC) (L-1, 0)                       partner rank ^ (1<<(p-1)), local index s ^ 1: every other element
//                                             of the partner's slab (packed by the sender)
// The slabs travel as the un-normalised w (NVLink, ncclSend/ncclRecv group); 1/beta is applied to the sum.
// Fused with the step's normalisation and the alpha dot like every other operator (op.cuh).
#include <algorithm>

#include "device_utils.cuh"
#include "op.cuh"

namespace cmb {

constexpr int kMaxFull = 6;  // rank-rank bonds (p <= 7)

struct HeisArgs {
  int Ll;             // local bits
  int nb;             // total number of bonds (for the diagonal)
  int n_local_bonds;  // bonds (i,i+1), i+1 < Ll
  int local_wrap;     // single rank + PBC: bond (Ll-1, 0) is local
  int aligned_uniform;  // rank-rank bonds whose two rank bits agree
  int n_full;           // rank-rank bonds whose rank bits differ: full partner slabs
  const double* full[kMaxFull];
  int has_straddle, rb0;  // bond (Ll-1, Ll)
  const double* half;
  int has_wrap, rt;  // bond (L-1, 0) across ranks
  const double* wrap;
  double J;
};

template <bool CPLX>
__global__ void __launch_bounds__(256)
heis_apply_kernel(HeisArgs a, const double* __restrict__ w, double* __restrict__ ucol, double* __restrict__ v,
                  double shr, double shi, StepScalars sc, double* partial, unsigned* ticket) {
  double inv;
  if (!step_prologue(sc, inv)) return;
  const long long dim = 1ll << a.Ll;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long top = 1ll << (a.Ll - 1);
  double d0 = 0.0, d1 = 0.0;
  auto ld = [&](const double* p, long long i, double& re, double& im) {
    if (CPLX) {
      const double2 z = reinterpret_cast<const double2*>(p)[i];
      re += z.x;
      im += z.y;
    } else {
      re += p[i];
    }
  };
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < dim; s += stride) {
    int aligned = a.aligned_uniform;
    double ar = 0.0, ai = 0.0;
    for (int b = 0; b < a.n_local_bonds; ++b) {
      if (((s >> b) ^ (s >> (b + 1))) & 1)
        ld(w, s ^ (3ll << b), ar, ai);
      else
        ++aligned;
    }
    if (a.local_wrap) {
      if (((s >> (a.Ll - 1)) ^ s) & 1)
        ld(w, s ^ (top | 1ll), ar, ai);
      else
        ++aligned;
    }
    for (int k = 0; k < a.n_full; ++k) ld(a.full[k], s, ar, ai);
    if (a.has_straddle) {
      if (int((s >> (a.Ll - 1)) & 1) != a.rb0)
        ld(a.half, s & (top - 1), ar, ai);
      else
        ++aligned;
    }
    if (a.has_wrap) {
      if (int(s & 1) != a.rt)
        ld(a.wrap, s >> 1, ar, ai);
      else
        ++aligned;
    }
    const double diag = 0.25 * a.J * double(2 * aligned - a.nb);
    if (CPLX) {
      const double2 wi = reinterpret_cast<const double2*>(w)[s];
      const double ur = wi.x * inv, ui = wi.y * inv;
      const double yr = (diag * wi.x + 0.5 * a.J * ar) * inv + (shr * ur - shi * ui);
      const double yi = (diag * wi.y + 0.5 * a.J * ai) * inv + (shr * ui + shi * ur);
      reinterpret_cast<double2*>(ucol)[s] = make_double2(ur, ui);
      reinterpret_cast<double2*>(v)[s] = make_double2(yr, yi);
      d0 += ur * yr + ui * yi;
      d1 += ur * yi - ui * yr;
    } else {
      const double wi = w[s];
      const double ui = wi * inv;
      const double y = (diag * wi + 0.5 * a.J * ar) * inv + shr * ui;
      ucol[s] = ui;
      v[s] = y;
      d0 = fma(ui, y, d0);
    }
  }
  grid_sum_finalize<CPLX ? 2 : 1>(d0, d1, partial, ticket, sc.alpha_slot);
}

// out[i] = w[2 i + parity] : the elements a wrap-bond partner needs
template <int ES>
__global__ void pack_parity_kernel(const double* __restrict__ w, long long half, int parity, double* __restrict__ out,
                                   const int* __restrict__ halt) {
  if (*halt) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
#pragma unroll
    for (int e = 0; e < ES; ++e) out[i * ES + e] = w[(2 * i + parity) * ES + e];
  }
}

// The remote bonds of one rank, in the order both partners enumerate them.
struct HeisRemote {
  int kind;     // 2 straddle, 3 rank-rank, 4 wrap
  int partner;  // rank
  bool needed;  // rank-rank bonds only exchange when the two rank bits differ
};

struct HeisPlan {
  int L = 0, p = 0, Ll = 0, rank = 0, nb = 0;
  bool pbc = false;
  std::vector<HeisRemote> remote;
  void build(int L_, bool pbc_, int P, int rank_) {
    L = L_;
    pbc = pbc_;
    rank = rank_;
    p = 0;
    while ((1 << p) < P) ++p;
    Ll = L - p;
    nb = (pbc && L > 2) ? L : L - 1;
    remote.clear();
    if (p == 0) return;
    remote.push_back({2, rank ^ 1, true});
    for (int i = Ll; i + 1 < L; ++i) {
      const int b = i - Ll;
      const bool differ = (((rank >> b) ^ (rank >> (b + 1))) & 1) != 0;
      remote.push_back({3, rank ^ (3 << b), differ});
    }
    if (nb == L) remote.push_back({4, rank ^ (1 << (p - 1)), true});
  }
};

struct HeisenbergOp : cmb_op {
  HeisPlan plan;
  double J = 1.0;
  // distributed state
  double* d_recv = nullptr;  // receive buffers, one region per needed remote bond
  double* d_pack = nullptr;  // packed parity half for the wrap bond
  std::vector<size_t> recv_off;
  ~HeisenbergOp() override {
    pool_free(ctx, d_recv);
    pool_free(ctx, d_pack);
  }

  HeisArgs base_args() const {
    HeisArgs a;
    memset(&a, 0, sizeof(a));
    a.Ll = plan.Ll;
    a.nb = plan.nb;
    a.n_local_bonds = plan.Ll - 1;
    a.local_wrap = (plan.p == 0 && plan.nb == plan.L) ? 1 : 0;
    a.J = J;
    return a;
  }

  int alloc_dist() {
    if (plan.p == 0) return CMB_OK;
    const size_t es = cplx ? 2 : 1;
    const size_t slab = size_t(n_local) * es, half = slab / 2;
    size_t tot = 0;
    recv_off.clear();
    for (auto& r : plan.remote) {
      recv_off.push_back(tot);
      if (!r.needed) continue;
      tot += (r.kind == 3) ? slab : half;
    }
    CMB_TRY(pool_alloc(ctx, &d_recv, sizeof(double) * std::max<size_t>(tot, 2)));
    CMB_TRY(pool_alloc(ctx, &d_pack, sizeof(double) * std::max<size_t>(half, 2)));
    return CMB_OK;
  }

  // fills the remote part of the kernel arguments from the receive buffers
  void remote_args(HeisArgs& a, const double* recv_base) const {
    for (size_t k = 0; k < plan.remote.size(); ++k) {
      const HeisRemote& r = plan.remote[k];
      const double* buf = recv_base + recv_off[k];
      if (r.kind == 2) {
        a.has_straddle = 1;
        a.rb0 = plan.rank & 1;
        a.half = buf;
      } else if (r.kind == 3) {
        if (r.needed)
          a.full[a.n_full++] = buf;
        else
          a.aligned_uniform++;
      } else {
        a.has_wrap = 1;
        a.rt = (plan.rank >> (plan.p - 1)) & 1;
        a.wrap = buf;
      }
    }
  }

  int launch(const HeisArgs& a, const double* w, double* ucol, double* v, double shr, double shi,
             const StepScalars& sc) {
    int grid = int(std::min<long long>((n_local + 255) / 256, (long long)ctx->num_sms * 8));
    LaunchScope ls(ctx, "heisenberg_mf");
    if (cplx)
      heis_apply_kernel<true><<<grid, 256, 0, ctx->stream>>>(a, w, ucol, v, shr, shi, sc, ctx->d_partial,
                                                             ctx->d_ticket + 1);
    else
      heis_apply_kernel<false><<<grid, 256, 0, ctx->stream>>>(a, w, ucol, v, shr, shi, sc, ctx->d_partial,
                                                              ctx->d_ticket + 1);
    CMB_CUDA(cudaGetLastError());
    return CMB_OK;
  }

  int pack_wrap(const double* w, int parity, double* out, const int* halt) {
    const long long half = n_local / 2;
    const int grid = int(std::max<long long>(1, std::min<long long>((half + 255) / 256, (long long)ctx->num_sms * 8)));
    LaunchScope ls(ctx, "halo_pack");
    if (cplx)
      pack_parity_kernel<2><<<grid, 256, 0, ctx->stream>>>(w, half, parity, out, halt);
    else
      pack_parity_kernel<1><<<grid, 256, 0, ctx->stream>>>(w, half, parity, out, halt);
    CMB_CUDA(cudaGetLastError());
    return CMB_OK;
  }

  int apply(const double* w, double* ucol, double* v, double shr, double shi, const StepScalars& sc) override {
    HeisArgs a = base_args();
    if (plan.p > 0) {
      const size_t es = cplx ? 2 : 1;
      const size_t slab = size_t(n_local) * es, half = slab / 2;
      // what each partner needs from me: straddle -> my half whose top local bit != my rank bit 0 (the partner's
      // rank bit 0); wrap -> my elements whose bit 0 != my top rank bit (packed)
      const int rb0 = plan.rank & 1, rt = (plan.rank >> (plan.p - 1)) & 1;
      bool has_wrap = false;
      for (auto& r : plan.remote) has_wrap |= (r.kind == 4);
      if (has_wrap) CMB_TRY(pack_wrap(w, 1 - rt, d_pack, sc.halt));
      auto chk = [&](int r, const char* what) -> int {
        if (r != 0) {
          set_error("%s failed: %s", what, ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "?");
          return CMB_ERR_NCCL;
        }
        return CMB_OK;
      };
      int rc = chk(ctx->nccl->GroupStart(), "ncclGroupStart");
      for (size_t k = 0; k < plan.remote.size() && rc == CMB_OK; ++k) {
        const HeisRemote& r = plan.remote[k];
        if (!r.needed) continue;
        const double* src = w;
        size_t count = slab;
        if (r.kind == 2) {
          src = w + size_t(1 - rb0) * half;
          count = half;
        } else if (r.kind == 4) {
          src = d_pack;
          count = half;
        }
        rc = chk(ctx->nccl->Send(src, count, kNcclFloat64, r.partner, ctx->nccl_comm, ctx->stream), "ncclSend");
        if (rc == CMB_OK)
          rc = chk(ctx->nccl->Recv(d_recv + recv_off[k], count, kNcclFloat64, r.partner, ctx->nccl_comm, ctx->stream),
                   "ncclRecv");
      }
      if (rc == CMB_OK)
        rc = chk(ctx->nccl->GroupEnd(), "ncclGroupEnd");
      else
        ctx->nccl->GroupEnd();
      CMB_TRY(rc);
      remote_args(a, d_recv);
    }
    return launch(a, w, ucol, v, shr, shi, sc);
  }
};

}  // namespace cmb

using namespace cmb;

extern "C" {

int cmb_op_heisenberg_create(cmb_ctx* ctx, cmb_dtype dtype, int L, double J, int pbc, cmb_op** out) {
  CMB_REQUIRE(ctx && out, "null argument");
  *out = nullptr;
  CMB_REQUIRE(dtype == CMB_F64 || dtype == CMB_C64, "dtype must be CMB_F64 or CMB_C64");
  CMB_REQUIRE(L >= 2 && L <= 40, "chain length out of range");
  CMB_CUDA(cudaSetDevice(ctx->device));
  HeisenbergOp* op = new (std::nothrow) HeisenbergOp();
  if (!op) return CMB_ERR_NOMEM;
  op->plan.build(L, pbc != 0, ctx->nranks, ctx->rank);
  if (op->plan.Ll < 2 || op->plan.p > kMaxFull + 1) {
    delete op;
    set_error("Heisenberg chain of %d sites cannot be split over %d ranks", L, ctx->nranks);
    return CMB_ERR_INVALID;
  }
  const int64_t n = int64_t(1) << L, nloc = int64_t(1) << op->plan.Ll;
  op->ctx = ctx;
  op->dtype = dtype;
  op->cplx = dtype == CMB_C64;
  op->n_global = n;
  op->row_begin = int64_t(ctx->rank) * nloc;
  op->n_local = nloc;
  op->family = "heisenberg_mf";
  op->J = J;
  int rc = op->alloc_dist();
  if (rc != CMB_OK) {
    delete op;
    return rc;
  }
  op->bytes = 2.0 * double(nloc) * (op->cplx ? 16.0 : 8.0);  // B_mf = 2 n s
  *out = op;
  return CMB_OK;
}

// Diagnostic / test entry: the distributed operator with P = 2^p VIRTUAL ranks on one GPU.  Every virtual rank
// runs the same kernel and the same packing as a real rank; the NCCL exchange is replaced by device copies
// between the virtual ranks' slabs.  x and y are full host vectors of 2^L elements.
int cmb_debug_heisenberg_virtual(cmb_ctx* ctx, cmb_dtype dtype, int L, double J, int pbc, int nranks, const void* x,
                                 void* y) {
  CMB_REQUIRE(ctx && x && y, "null argument");
  CMB_REQUIRE(nranks >= 1 && (nranks & (nranks - 1)) == 0, "nranks must be a power of two");
  CMB_CUDA(cudaSetDevice(ctx->device));
  const bool cplx = dtype == CMB_C64;
  const size_t es = cplx ? 2 : 1;
  int p = 0;
  while ((1 << p) < nranks) ++p;
  const int Ll = L - p;
  CMB_REQUIRE(Ll >= 2 && L <= 28, "bad chain length for the virtual-rank test");
  const size_t nloc = size_t(1) << Ll, slab = nloc * es, half = slab / 2, n = size_t(1) << L;
  double *d_x = nullptr, *d_u = nullptr, *d_v = nullptr, *d_recv = nullptr, *d_pack = nullptr, *d_sc = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_x);
    cudaFree(d_u);
    cudaFree(d_v);
    cudaFree(d_recv);
    cudaFree(d_pack);
    cudaFree(d_sc);
  };
  int rc = [&]() -> int {
    CMB_CUDA(cudaMalloc(&d_x, sizeof(double) * n * es));
    CMB_CUDA(cudaMalloc(&d_u, sizeof(double) * slab));
    CMB_CUDA(cudaMalloc(&d_v, sizeof(double) * slab));
    CMB_CUDA(cudaMalloc(&d_recv, sizeof(double) * slab * (kMaxFull + 1)));
    CMB_CUDA(cudaMalloc(&d_pack, sizeof(double) * std::max<size_t>(half, 2)));
    CMB_CUDA(cudaMalloc(&d_sc, sizeof(double) * 8));
    CMB_CUDA(cudaMemcpyAsync(d_x, x, sizeof(double) * n * es, cudaMemcpyHostToDevice, ctx->stream));
    for (int r = 0; r < nranks; ++r) {
      HeisenbergOp op;
      op.ctx = ctx;
      op.cplx = cplx;
      op.n_local = int64_t(nloc);
      op.J = J;
      op.plan.build(L, pbc != 0, nranks, r);
      op.recv_off.clear();
      size_t tot = 0;
      for (auto& rem : op.plan.remote) {
        op.recv_off.push_back(tot);
        if (rem.needed) tot += (rem.kind == 3) ? slab : half;
      }
      CMB_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(double) * 8, ctx->stream));
      const double one = 1.0;
      CMB_CUDA(cudaMemcpyAsync(d_sc, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      StepScalars sc;
      sc.nrm2 = d_sc;
      sc.threshold = -1.0;
      sc.halt = reinterpret_cast<int*>(d_sc + 1);
      sc.beta_slot = d_sc + 2;
      sc.alpha_slot = d_sc + 4;
      // "receive": what the partner would have sent me
      for (size_t k = 0; k < op.plan.remote.size(); ++k) {
        const HeisRemote& rem = op.plan.remote[k];
        if (!rem.needed) continue;
        const double* pw = d_x + size_t(rem.partner) * slab;  // the partner's slab
        const int prb0 = rem.partner & 1, prt = (rem.partner >> (p - 1)) & 1;
        if (rem.kind == 2) {
          CMB_CUDA(cudaMemcpyAsync(d_recv + op.recv_off[k], pw + size_t(1 - prb0) * half, sizeof(double) * half,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
        } else if (rem.kind == 3) {
          CMB_CUDA(cudaMemcpyAsync(d_recv + op.recv_off[k], pw, sizeof(double) * slab, cudaMemcpyDeviceToDevice,
                                   ctx->stream));
        } else {
          CMB_TRY(op.pack_wrap(pw, 1 - prt, d_pack, sc.halt));
          CMB_CUDA(cudaMemcpyAsync(d_recv + op.recv_off[k], d_pack, sizeof(double) * half, cudaMemcpyDeviceToDevice,
                                   ctx->stream));
        }
      }
      HeisArgs a = op.base_args();
      op.remote_args(a, d_recv);
      CMB_TRY(op.launch(a, d_x + size_t(r) * slab, d_u, d_v, 0.0, 0.0, sc));
      CMB_CUDA(cudaMemcpyAsync(static_cast<double*>(y) + size_t(r) * slab, d_v, sizeof(double) * slab,
                               cudaMemcpyDeviceToHost, ctx->stream));
      CMB_CUDA(cudaStreamSynchronize(ctx->stream));
      op.ctx = nullptr;  // stack object: nothing to free through the pool
    }
    return CMB_OK;
  }();
  cleanup();
  return rc;
}

}  // extern "C"
