// ctx.cu — context (one GPU / one rank), error strings, NCCL loader, timers, profiling brackets.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <algorithm>

#include "cmpt_b200_debug.h"
#include "common.cuh"

namespace cmb {

VGroup* vgroup_of(cmb_vgroup* vg);  // vgroup.cu

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int nccl_load(NcclApi** out) {
  static NcclApi api;
  static int state = 0;  // 0 = not tried, 1 = ok, -1 = failed
  if (state == 0) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      state = -1;
    } else {
      auto sym = [&](const char* s) { return dlsym(api.handle, s); };
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
      api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
      api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
      api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
      api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
      state = (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Send && api.Recv &&
               api.GroupStart && api.GroupEnd)
                  ? 1
                  : -1;
    }
  }
  if (state != 1) {
    set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
    return CMB_ERR_NCCL;
  }
  *out = &api;
  return CMB_OK;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

LaunchScope::LaunchScope(cmb_ctx* c, const char* fam) : ctx(c), family(fam) {
  ctx->launches++;
  if (ctx->profiling) {
    auto get = [&]() {
      cudaEvent_t e = nullptr;
      if (!ctx->ev_pool.empty()) {
        e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
      } else {
        cudaEventCreate(&e);
      }
      return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, ctx->stream);
  }
}
LaunchScope::~LaunchScope() {
  if (a) {
    cudaEventRecord(b, ctx->stream);
    ctx->pending.push_back({family, a, b});
  }
}

int resolve_profile(cmb_ctx* ctx) {
  if (ctx->pending.empty()) return CMB_OK;
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& p : ctx->pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    auto& e = ctx->prof[p.family];
    e.ms += ms;
    e.launches++;
    ctx->ev_pool.push_back(p.a);
    ctx->ev_pool.push_back(p.b);
  }
  ctx->pending.clear();
  return CMB_OK;
}

int allreduce_sum_f64(cmb_ctx* ctx, double* p, size_t count) {
  if (ctx->nranks == 1) return CMB_OK;
  if (ctx->vgroup) return vgroup_allreduce_f64(ctx, p, count);
  LaunchScope ls(ctx, "nccl_allreduce");
  int r = ctx->nccl->AllReduce(p, p, count, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
  if (r != 0) {
    set_error("ncclAllReduce failed: %s", ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "?");
    return CMB_ERR_NCCL;
  }
  return CMB_OK;
}

int allreduce_min_u64(cmb_ctx* ctx, unsigned long long* p, size_t count) {
  if (ctx->nranks == 1) return CMB_OK;
  if (ctx->vgroup) return vgroup_allreduce_min_u64(ctx, p, count);
  int r = ctx->nccl->AllReduce(p, p, count, kNcclUint64, kNcclMin, ctx->nccl_comm, ctx->stream);
  if (r != 0) {
    set_error("ncclAllReduce(min) failed: %s", ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "?");
    return CMB_ERR_NCCL;
  }
  return CMB_OK;
}

// ---- peer-memory mailboxes -------------------------------------------------------------------------------
// Layout of one rank's mailbox: data [kMailSlots][P][kMailStride] doubles, then flags [kMailSlots][P] u64.
static size_t mail_data_doubles(int P) { return size_t(kMailSlots) * P * kMailStride; }

MailPush mail_next_push(cmb_ctx* ctx) {
  MailPush m;
  m.P = ctx->nranks;
  m.rank = ctx->rank;
  m.seq = ++ctx->mail_seq;
  const size_t slot = size_t(m.seq % kMailSlots);
  for (int q = 0; q < kMaxPeers; ++q) {
    m.data[q] = nullptr;
    m.flag[q] = nullptr;
  }
  for (int q = 0; q < ctx->nranks; ++q) {
    m.data[q] = ctx->mail_data[q] + slot * ctx->nranks * kMailStride;
    m.flag[q] = ctx->mail_flag[q] + slot * ctx->nranks;
  }
  return m;
}

MailPull mail_pull_of(cmb_ctx* ctx, unsigned long long seq, double* writeback) {
  MailPull m;
  m.P = ctx->nranks;
  m.seq = seq;
  const size_t slot = size_t(seq % kMailSlots);
  m.data = ctx->mail_data[ctx->rank] + slot * ctx->nranks * kMailStride;
  m.flag = ctx->mail_flag[ctx->rank] + slot * ctx->nranks;
  m.writeback = writeback;
  m.error = ctx->d_mail_error;
  m.timeout = ctx->spin_timeout;
  return m;
}

int rank_barrier(cmb_ctx* ctx) {
  if (ctx->nranks == 1) return CMB_OK;
  if (ctx->vgroup) return vgroup_barrier(ctx);
  CMB_TRY(allreduce_sum_f64(ctx, ctx->d_partial, 1));  // scratch word; its value is never read
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CMB_OK;
}

// Collective: every rank reports its local outcome and all of them leave with an error if any of them has one, so that
// no rank goes on to a collective step (or to a solve) that a failed peer will never join.
int agree_status(cmb_ctx* ctx, int rc, const char* what) {
  if (ctx->nranks == 1) return rc;
  const std::string mine = rc != CMB_OK ? std::string(cmb_last_error()) : std::string();
  double flag = rc == CMB_OK ? 0.0 : 1.0;
  double* word = ctx->d_partial;  // scratch
  if (cudaMemcpyAsync(word, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    set_error("%s: agreement between the ranks failed: %s", what, cudaGetErrorString(cudaGetLastError()));
    return CMB_ERR_CUDA;
  }
  CMB_TRY(allreduce_sum_f64(ctx, word, 1));
  double failed = 0.0;
  CMB_TRY(d2h_sync(ctx, &failed, word, sizeof(double)));
  if (rc != CMB_OK) {
    set_error("%s", mine.c_str());
    return rc;
  }
  if (failed > 0.0) {
    set_error("%s failed on %d other rank(s)", what, int(failed));
    return CMB_ERR_INVALID;
  }
  return CMB_OK;
}

int alltoallv_i32(cmb_ctx* ctx, const int32_t* d_send, const int64_t* send_off, int32_t* d_recv, const int64_t* recv_off) {
  if (ctx->nranks == 1) return CMB_OK;
  if (ctx->vgroup) return vgroup_alltoallv_i32(ctx, d_send, send_off, d_recv, recv_off);
  auto chk = [&](int r, const char* what) {
    if (r != 0) {
      set_error("%s failed: %s", what, ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "?");
      return int(CMB_ERR_NCCL);
    }
    return int(CMB_OK);
  };
  int rc = chk(ctx->nccl->GroupStart(), "ncclGroupStart");
  for (int q = 0; q < ctx->nranks && rc == CMB_OK; ++q) {
    if (q == ctx->rank) continue;
    if (send_off[q + 1] > send_off[q])
      rc = chk(ctx->nccl->Send(d_send + send_off[q], size_t(send_off[q + 1] - send_off[q]), kNcclInt32, q, ctx->nccl_comm,
                               ctx->stream), "ncclSend");
    if (rc == CMB_OK && recv_off[q + 1] > recv_off[q])
      rc = chk(ctx->nccl->Recv(d_recv + recv_off[q], size_t(recv_off[q + 1] - recv_off[q]), kNcclInt32, q, ctx->nccl_comm,
                               ctx->stream), "ncclRecv");
  }
  if (rc == CMB_OK) rc = chk(ctx->nccl->GroupEnd(), "ncclGroupEnd");
  else ctx->nccl->GroupEnd();
  return rc;
}

int check_peer_wait(cmb_ctx* ctx) {
  if (ctx->dead) {
    set_error("this context is unusable: an earlier wait for a peer rank timed out and the ranks' sequence numbers no "
              "longer agree; destroy the contexts of all ranks and create new ones");
    return CMB_ERR_NCCL;
  }
  if (!ctx->mail_ok) return CMB_OK;
  int err = 0;
  CMB_TRY(d2h_sync(ctx, &err, ctx->d_mail_error, sizeof(int)));
  if (!err) return CMB_OK;
  // the flag is cleared so that the message appears once; the context stays marked dead (see above)
  CMB_CUDA(cudaMemsetAsync(ctx->d_mail_error, 0, sizeof(int), ctx->stream));
  CMB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->dead = true;
  set_error("a peer rank never published its Gram-Schmidt partials or halo values: the in-kernel wait timed out after "
            "%.1f s (CMPT_B200_SPIN_TIMEOUT_S / cmb_ctx_set_spin_timeout change the limit; replaying profilers and "
            "debuggers are not compatible with multi-rank runs)", double(ctx->spin_timeout) / 1.9e9);
  return CMB_ERR_NCCL;
}

// Collective: every rank passes the base of a cudaMalloc'ed buffer; CUDA IPC handles travel through NCCL (allgather
// of raw bytes) and every peer's buffer gets mapped here (mapped[q], mapped[rank] = base).  All-or-nothing: returns
// true on every rank or false on every rank (then nothing stays mapped).
bool ipc_share(cmb_ctx* c, void* base, void** mapped) {
  const int P = c->nranks;
  if (c->vgroup) return vgroup_share(c, base, mapped);
  for (int q = 0; q < P; ++q) mapped[q] = nullptr;
  if (P < 2 || P > kMaxPeers || !c->nccl || !c->nccl->AllGather) return false;
  cudaIpcMemHandle_t mine;
  const bool ok = base && cudaIpcGetMemHandle(&mine, base) == cudaSuccess;
  if (!ok) cudaGetLastError();
  constexpr size_t kRec = 80;
  static_assert(sizeof(cudaIpcMemHandle_t) <= kRec - 8, "IPC handle larger than expected");
  unsigned char rec[kRec] = {0};
  rec[0] = ok ? 1 : 0;
  if (ok) memcpy(rec + 8, &mine, sizeof(mine));
  unsigned char *d_send = nullptr, *d_recv = nullptr;
  std::vector<unsigned char> all(kRec * P, 0);
  cudaMalloc(&d_send, kRec);
  cudaMalloc(&d_recv, kRec * P);
  cudaMemcpyAsync(d_send, rec, kRec, cudaMemcpyHostToDevice, c->stream);
  int nr = c->nccl->AllGather(d_send, d_recv, kRec, kNcclInt8, c->nccl_comm, c->stream);
  cudaMemcpyAsync(all.data(), d_recv, kRec * P, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t se = cudaStreamSynchronize(c->stream);
  cudaFree(d_send);
  cudaFree(d_recv);
  bool all_ok = (nr == 0) && (se == cudaSuccess);
  for (int q = 0; q < P && all_ok; ++q) all_ok = all[kRec * q] == 1;
  if (all_ok) {
    for (int q = 0; q < P; ++q) {
      void* p = base;
      if (q != c->rank) {
        cudaIpcMemHandle_t h;
        memcpy(&h, all.data() + kRec * q + 8, sizeof(h));
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          all_ok = false;
          break;
        }
      }
      mapped[q] = p;
    }
  }
  // every rank must agree on whether the mapping is usable (a single failure disables it everywhere)
  double* d_flag = nullptr;
  cudaMalloc(&d_flag, sizeof(double));
  const double mineok = all_ok ? 0.0 : 1.0;
  cudaMemcpyAsync(d_flag, &mineok, sizeof(double), cudaMemcpyHostToDevice, c->stream);
  c->nccl->AllReduce(d_flag, d_flag, 1, kNcclFloat64, kNcclSum, c->nccl_comm, c->stream);
  double failures = 1.0;
  cudaMemcpyAsync(&failures, d_flag, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  cudaStreamSynchronize(c->stream);
  cudaFree(d_flag);
  cudaGetLastError();
  if (failures != 0.0) {
    ipc_unshare(c, mapped);
    return false;
  }
  return true;
}

void ipc_unshare(cmb_ctx* c, void** mapped) {
  for (int q = 0; q < c->nranks; ++q) {
    if (q != c->rank && mapped[q] && !c->vgroup) cudaIpcCloseMemHandle(mapped[q]);
    mapped[q] = nullptr;
  }
  cudaGetLastError();
}

// Allocate this rank's mailbox and map the peers' mailboxes.  Failure is not fatal: the context then keeps using
// NCCL allreduce for the coefficients.
static int mail_setup(cmb_ctx* c) {
  if (c->nranks < 2 || c->nranks > kMaxPeers || (!c->vgroup && !c->nccl->AllGather)) return CMB_OK;
  if (getenv("CMPT_B200_NO_MAILBOX")) return CMB_OK;
  const int P = c->nranks;
  const size_t bytes = mail_data_doubles(P) * sizeof(double) + size_t(kMailSlots) * P * sizeof(unsigned long long);
  void* base = nullptr;
  if (cudaMalloc(&base, bytes) != cudaSuccess) {
    cudaGetLastError();
    base = nullptr;
  } else {
    cudaMemsetAsync(base, 0, bytes, c->stream);
  }
  cudaMalloc(&c->d_mail_error, sizeof(int));
  cudaMemsetAsync(c->d_mail_error, 0, sizeof(int), c->stream);
  cudaStreamSynchronize(c->stream);
  void* mapped[kMaxPeers];
  c->mail_ok = ipc_share(c, base, mapped);
  if (!c->mail_ok) {
    for (int q = 0; q < P; ++q) c->mail_data[q] = nullptr, c->mail_flag[q] = nullptr;
    if (base) cudaFree(base);
  } else {
    for (int q = 0; q < P; ++q) {
      c->mail_data[q] = static_cast<double*>(mapped[q]);
      c->mail_flag[q] = reinterpret_cast<unsigned long long*>(static_cast<double*>(mapped[q]) + mail_data_doubles(P));
    }
  }
  return CMB_OK;
}

static void mail_teardown(cmb_ctx* c) {
  if (!c->mail_ok) return;
  for (int q = 0; q < c->nranks; ++q)
    if (q != c->rank && c->mail_data[q] && !c->vgroup) cudaIpcCloseMemHandle(c->mail_data[q]);
  cudaFree(c->mail_data[c->rank]);
  cudaFree(c->d_mail_error);
  c->mail_ok = false;
}

static int ctx_init_common(cmb_ctx* c, int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("no CUDA device is visible: libcmpt_b200 has no CPU fallback");
    return CMB_ERR_NO_DEVICE;
  }
  CMB_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  c->device = device;
  CMB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CMB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return CMB_ERR_UNSUPPORTED;
  }
  c->num_sms = prop.multiProcessorCount;
  {
    double seconds = 30.0;
    if (const char* t = getenv("CMPT_B200_SPIN_TIMEOUT_S")) seconds = std::max(0.1, atof(t));
    c->spin_timeout = (long long)(seconds * double(prop.clockRate) * 1e3);
    if (const char* g = getenv("CMPT_B200_NORM_GUARD")) c->norm_guard = atof(g);
  }
  {
    // keep freed pool memory cached: operator (re)builds then never reach cudaMalloc/cudaFree
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
  }
  CMB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CMB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CMB_CUDA(cudaEventCreate(&c->t0));
  CMB_CUDA(cudaEventCreate(&c->t1));
  CMB_CUDA(cudaMalloc(&c->d_partial, sizeof(double) * size_t(kMaxGrid) * kPartialStride));
  CMB_CUDA(cudaMalloc(&c->d_ticket, sizeof(unsigned) * 16));
  CMB_CUDA(cudaMemsetAsync(c->d_ticket, 0, sizeof(unsigned) * 16, c->stream));
  if (!get_encode_tiled()) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return CMB_ERR_CUDA;
  }
  return CMB_OK;
}

}  // namespace cmb

using namespace cmb;

extern "C" {

const char* cmb_version(void) { return "cmpt_b200 0.1 (sm_100a)"; }
const char* cmb_last_error(void) { return get_error(); }

int cmb_device_count(int* count) {
  CMB_REQUIRE(count, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return CMB_OK;
}

int cmb_ctx_create(int device, cmb_ctx** out) {
  CMB_REQUIRE(out, "null argument");
  *out = nullptr;
  cmb_ctx* c = new (std::nothrow) cmb_ctx();
  if (!c) return CMB_ERR_NOMEM;
  int r = ctx_init_common(c, device);
  if (r != CMB_OK) {
    delete c;
    return r;
  }
  *out = c;
  return CMB_OK;
}

int cmb_nccl_unique_id(void* id128) {
  CMB_REQUIRE(id128, "null argument");
  NcclApi* api = nullptr;
  CMB_TRY(nccl_load(&api));
  NcclId id;
  int r = api->GetUniqueId(&id);
  if (r != 0) {
    set_error("ncclGetUniqueId failed (%d)", r);
    return CMB_ERR_NCCL;
  }
  memcpy(id128, &id, sizeof(id));
  return CMB_OK;
}

int cmb_ctx_create_dist(int device, int rank, int nranks, const void* nccl_id, cmb_ctx** out) {
  CMB_REQUIRE(out, "null argument");
  *out = nullptr;
  CMB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  CMB_REQUIRE((nranks & (nranks - 1)) == 0, "nranks must be a power of two");
  cmb_ctx* c = new (std::nothrow) cmb_ctx();
  if (!c) return CMB_ERR_NOMEM;
  int r = ctx_init_common(c, device);
  if (r != CMB_OK) {
    delete c;
    return r;
  }
  c->rank = rank;
  c->nranks = nranks;
  if (nranks > 1) {
    if (!nccl_id) {
      set_error("nccl_id is required when nranks > 1");
      delete c;
      return CMB_ERR_INVALID;
    }
    r = nccl_load(&c->nccl);
    if (r != CMB_OK) {
      delete c;
      return r;
    }
    NcclId id;
    memcpy(&id, nccl_id, sizeof(id));
    int nr = c->nccl->CommInitRank(&c->nccl_comm, nranks, id, rank);
    if (nr != 0) {
      set_error("ncclCommInitRank failed: %s", c->nccl->GetErrorString ? c->nccl->GetErrorString(nr) : "?");
      delete c;
      return CMB_ERR_NCCL;
    }
    mail_setup(c);
  }
  *out = c;
  return CMB_OK;
}

int cmb_ctx_destroy(cmb_ctx* c) {
  if (!c) return CMB_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->vgroup && !c->mail_ok && c->nranks > 1) rank_barrier(c);
  if (c->mail_ok) {
    // unmap the peers' mailboxes, wait until everybody has done so, then free the own one
    for (int q = 0; q < c->nranks; ++q)
      if (q != c->rank && c->mail_data[q]) {
        if (!c->vgroup) cudaIpcCloseMemHandle(c->mail_data[q]);
        c->mail_data[q] = nullptr;
      }
    rank_barrier(c);
    mail_teardown(c);
  }
  if (c->nccl_comm && c->nccl) c->nccl->CommDestroy(c->nccl_comm);
  for (auto& p : c->pending) {
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  if (c->d_partial) cudaFree(c->d_partial);
  if (c->d_ticket) cudaFree(c->d_ticket);
  if (c->d_flush) cudaFree(c->d_flush);
  if (c->t0) cudaEventDestroy(c->t0);
  if (c->t1) cudaEventDestroy(c->t1);
  for (void* p : c->graveyard) cudaFree(p);
  for (void* p : c->host_graveyard) cudaFreeHost(p);
  if (c->mempool) {
    cudaStreamSynchronize(c->stream);
    cudaMemPoolDestroy(c->mempool);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->vgroup) vgroup_detach(c->vgroup);
  delete c;
  return CMB_OK;
}

// Virtual rank `rank` of the group (declared in cmpt_b200_debug.h).  Collective: call it from nranks host threads.
int cmb_ctx_create_virtual(cmb_vgroup* vg, int rank, cmb_ctx** out) {
  CMB_REQUIRE(vg && out, "null argument");
  *out = nullptr;
  VGroup* g = vgroup_of(vg);
  CMB_REQUIRE(rank >= 0 && rank < vgroup_size(g), "bad virtual rank");
  cmb_ctx* c = new (std::nothrow) cmb_ctx();
  if (!c) return CMB_ERR_NOMEM;
  int r = ctx_init_common(c, vgroup_device(g));
  if (r == CMB_OK) r = vgroup_attach(g, c, rank);
  if (r != CMB_OK) {
    if (c->vgroup) vgroup_detach(c->vgroup);
    delete c;
    return r;
  }
  if (c->nranks > 1) mail_setup(c);
  if (c->nranks > 1 && !c->mail_ok) {
    set_error("virtual rank %d: the peer-memory mailboxes could not be set up", rank);
    cmb_ctx_destroy(c);
    return CMB_ERR_NCCL;
  }
  *out = c;
  return CMB_OK;
}

int cmb_ctx_set_spin_timeout(cmb_ctx* c, double seconds) {
  CMB_REQUIRE(c && seconds > 0.0, "bad argument");
  cudaDeviceProp prop;
  CMB_CUDA(cudaGetDeviceProperties(&prop, c->device));
  c->spin_timeout = (long long)(seconds * double(prop.clockRate) * 1e3);
  return CMB_OK;
}

int cmb_ctx_rank(const cmb_ctx* c) { return c ? c->rank : 0; }
int cmb_ctx_nranks(const cmb_ctx* c) { return c ? c->nranks : 1; }

int cmb_ctx_sync(cmb_ctx* c) {
  CMB_REQUIRE(c, "null context");
  CMB_CUDA(cudaSetDevice(c->device));
  CMB_CUDA(cudaStreamSynchronize(c->stream));
  return CMB_OK;
}

int cmb_ctx_timer_start(cmb_ctx* c) {
  CMB_REQUIRE(c, "null context");
  CMB_CUDA(cudaSetDevice(c->device));
  CMB_CUDA(cudaEventRecord(c->t0, c->stream));
  return CMB_OK;
}

int cmb_ctx_timer_stop(cmb_ctx* c, double* ms) {
  CMB_REQUIRE(c && ms, "null argument");
  CMB_CUDA(cudaSetDevice(c->device));
  CMB_CUDA(cudaEventRecord(c->t1, c->stream));
  CMB_CUDA(cudaEventSynchronize(c->t1));
  float f = 0.f;
  CMB_CUDA(cudaEventElapsedTime(&f, c->t0, c->t1));
  *ms = f;
  return CMB_OK;
}

uint64_t cmb_ctx_launch_count(const cmb_ctx* c) { return c ? c->launches : 0; }

int cmb_ctx_profile(cmb_ctx* c, int enable) {
  CMB_REQUIRE(c, "null context");
  CMB_TRY(resolve_profile(c));
  c->profiling = enable != 0;
  if (enable) c->prof.clear();
  return CMB_OK;
}

int cmb_ctx_profile_get(cmb_ctx* c, const char* family, double* total_ms, uint64_t* launches) {
  CMB_REQUIRE(c && family, "null argument");
  CMB_TRY(resolve_profile(c));
  auto it = c->prof.find(family);
  if (total_ms) *total_ms = it == c->prof.end() ? 0.0 : it->second.ms;
  if (launches) *launches = it == c->prof.end() ? 0 : it->second.launches;
  return CMB_OK;
}

int cmb_ctx_flush_l2(cmb_ctx* c) {
  CMB_REQUIRE(c, "null context");
  CMB_CUDA(cudaSetDevice(c->device));
  if (!c->d_flush) {
    c->flush_bytes = size_t(256) << 20;  // 256 MiB > 126 MB L2
    CMB_CUDA(cudaMalloc(&c->d_flush, c->flush_bytes));
  }
  CMB_CUDA(cudaMemsetAsync(c->d_flush, 0, c->flush_bytes, c->stream));
  return CMB_OK;
}

}  // extern "C"
