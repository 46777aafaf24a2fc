// vgroup.cu — virtual ranks: P row-partitioned contexts on ONE device inside ONE process.
//
// Purpose: the code that only runs with more than one rank — the peer-memory mailboxes of the Gram-Schmidt passes,
// the halo push fused into the SpMV, the all-push mode, the Pythagorean beta, the slab exchange of the matrix-free
// Heisenberg apply — must be testable on a box with a single GPU (tests/test_virtual_ranks.py).  A virtual rank is
// an ordinary cmb_ctx whose
//   * compute stream lives in a CUDA green context that owns a disjoint share of the SMs (so the kernels of the P
//     ranks are co-resident like kernels on P GPUs, and a kernel that spins on a peer's flag cannot starve the peer),
//   * peer buffers are plain device pointers handed over through the group (instead of CUDA-IPC handles through NCCL),
//   * host-side collectives (bootstrap all-gathers, the few allreduces outside the step chain) are rendezvous of the
//     P host threads that drive the ranks (instead of NCCL calls).
// Everything on the per-step path — kernels, flags, sequence numbers, double buffering — is the code real ranks run.
// Each virtual rank must be driven by its own host thread: the collectives block until all P ranks arrive.
#include <chrono>
#include <condition_variable>
#include <mutex>

#include "common.cuh"
#include "cmpt_b200_debug.h"

namespace cmb {

struct VGroup {
  int device = 0, P = 1;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t generation = 0;
  bool broken = false;
  double timeout_s = 120.0;
  std::vector<const void*> slot;               // one pointer per rank (valid between two barriers)
  std::vector<std::vector<int64_t>> offsets;   // one offset table per rank
  // green contexts (optional)
  bool green = false;
  std::vector<CUgreenCtx> gctx;
  std::vector<int> sms;
  int attached = 0;
};

namespace {

struct GreenApi {
  bool ok = false;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                        unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
};

GreenApi& green_api() {
  static GreenApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    auto get = [](const char* name) -> void* {
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      return p;
    };
    api.DeviceGet = reinterpret_cast<decltype(api.DeviceGet)>(get("cuDeviceGet"));
    api.DeviceGetDevResource = reinterpret_cast<decltype(api.DeviceGetDevResource)>(get("cuDeviceGetDevResource"));
    api.DevSmResourceSplitByCount = reinterpret_cast<decltype(api.DevSmResourceSplitByCount)>(get("cuDevSmResourceSplitByCount"));
    api.DevResourceGenerateDesc = reinterpret_cast<decltype(api.DevResourceGenerateDesc)>(get("cuDevResourceGenerateDesc"));
    api.GreenCtxCreate = reinterpret_cast<decltype(api.GreenCtxCreate)>(get("cuGreenCtxCreate"));
    api.GreenCtxStreamCreate = reinterpret_cast<decltype(api.GreenCtxStreamCreate)>(get("cuGreenCtxStreamCreate"));
    api.GreenCtxDestroy = reinterpret_cast<decltype(api.GreenCtxDestroy)>(get("cuGreenCtxDestroy"));
    api.ok = api.DeviceGet && api.DeviceGetDevResource && api.DevSmResourceSplitByCount && api.DevResourceGenerateDesc &&
             api.GreenCtxCreate && api.GreenCtxStreamCreate && api.GreenCtxDestroy;
  }
  return api;
}

// Splits the SMs of the device into P equal groups and creates one green context per group.
void make_green_contexts(VGroup* g, int total_sms) {
  g->sms.assign(g->P, std::max(1, total_sms / g->P));
  if (getenv("CMPT_B200_NO_GREEN_CTX")) return;
  GreenApi& api = green_api();
  if (!api.ok) return;
  CUdevice dev;
  if (api.DeviceGet(&dev, g->device) != CUDA_SUCCESS) return;
  CUdevResource all;
  if (api.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return;
  // groups of a multiple of 8 SMs (the granularity of compute capability 9.0+)
  unsigned want = unsigned(total_sms / g->P) / 8u * 8u;
  if (want < 8u) return;
  std::vector<CUdevResource> parts(g->P);
  unsigned ngroups = unsigned(g->P);
  CUdevResource rest;
  if (api.DevSmResourceSplitByCount(parts.data(), &ngroups, &all, &rest, 0, want) != CUDA_SUCCESS || int(ngroups) < g->P) return;
  std::vector<CUgreenCtx> ctxs;
  for (int q = 0; q < g->P; ++q) {
    CUdevResourceDesc desc;
    CUgreenCtx gc = nullptr;
    if (api.DevResourceGenerateDesc(&desc, &parts[q], 1) != CUDA_SUCCESS ||
        api.GreenCtxCreate(&gc, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
      for (auto c : ctxs) api.GreenCtxDestroy(c);
      return;
    }
    ctxs.push_back(gc);
    g->sms[q] = int(parts[q].sm.smCount);
  }
  g->gctx = ctxs;
  g->green = true;
}

// Generation barrier over the P host threads.  Returns false when a rank never arrives (or the group is broken).
bool vg_barrier(VGroup* g) {
  std::unique_lock<std::mutex> lk(g->mu);
  if (g->broken) return false;
  const uint64_t gen = g->generation;
  if (++g->arrived == g->P) {
    g->arrived = 0;
    ++g->generation;
    g->cv.notify_all();
    return true;
  }
  const bool ok = g->cv.wait_for(lk, std::chrono::duration<double>(g->timeout_s), [&] { return g->generation != gen || g->broken; });
  if (!ok || g->broken) {
    g->broken = true;
    g->cv.notify_all();
    return false;
  }
  return true;
}

int vg_fail(const char* what) {
  set_error("virtual-rank group: %s (a rank did not arrive at a collective within the time limit; every virtual rank "
            "needs its own host thread making the same sequence of calls)", what);
  return CMB_ERR_NCCL;
}

}  // namespace

int vgroup_allreduce_f64(cmb_ctx* c, double* p, size_t count) {
  VGroup* g = c->vgroup;
  CMB_CUDA(cudaStreamSynchronize(c->stream));
  g->slot[c->rank] = p;
  if (!vg_barrier(g)) return vg_fail("allreduce");
  std::vector<double> acc(count, 0.0), tmp(count);
  int rc = CMB_OK;
  for (int q = 0; q < g->P && rc == CMB_OK; ++q) {  // rank order: identical bits on every rank
    if (d2h_sync(c, tmp.data(), g->slot[q], sizeof(double) * count) != CMB_OK) {
      set_error("virtual allreduce: %s", cudaGetErrorString(cudaGetLastError()));
      rc = CMB_ERR_CUDA;
    }
    for (size_t i = 0; i < count; ++i) acc[i] += tmp[i];
  }
  if (!vg_barrier(g)) return vg_fail("allreduce");  // everybody has read the inputs
  CMB_TRY(rc);
  CMB_CUDA(cudaMemcpyAsync(p, acc.data(), sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
  CMB_CUDA(cudaStreamSynchronize(c->stream));
  return CMB_OK;
}

int vgroup_allreduce_min_u64(cmb_ctx* c, unsigned long long* p, size_t count) {
  VGroup* g = c->vgroup;
  CMB_CUDA(cudaStreamSynchronize(c->stream));
  g->slot[c->rank] = p;
  if (!vg_barrier(g)) return vg_fail("allreduce(min)");
  std::vector<unsigned long long> acc(count, ~0ull), tmp(count);
  int rc = CMB_OK;
  for (int q = 0; q < g->P && rc == CMB_OK; ++q) {
    if (d2h_sync(c, tmp.data(), g->slot[q], sizeof(unsigned long long) * count) != CMB_OK) {
      set_error("virtual allreduce(min): %s", cudaGetErrorString(cudaGetLastError()));
      rc = CMB_ERR_CUDA;
    }
    for (size_t i = 0; i < count; ++i) acc[i] = std::min(acc[i], tmp[i]);
  }
  if (!vg_barrier(g)) return vg_fail("allreduce(min)");
  CMB_TRY(rc);
  CMB_CUDA(cudaMemcpyAsync(p, acc.data(), sizeof(unsigned long long) * count, cudaMemcpyHostToDevice, c->stream));
  CMB_CUDA(cudaStreamSynchronize(c->stream));
  return CMB_OK;
}

// The virtual counterpart of ipc_share: same process, same device, so a peer's buffer is simply its pointer.
bool vgroup_share(cmb_ctx* c, void* base, void** mapped) {
  VGroup* g = c->vgroup;
  for (int q = 0; q < g->P; ++q) mapped[q] = nullptr;
  cudaStreamSynchronize(c->stream);
  g->slot[c->rank] = base;
  if (!vg_barrier(g)) return false;
  bool ok = true;
  for (int q = 0; q < g->P; ++q) {
    mapped[q] = const_cast<void*>(g->slot[q]);
    ok = ok && mapped[q] != nullptr;
  }
  if (!vg_barrier(g)) return false;
  if (!ok)
    for (int q = 0; q < g->P; ++q) mapped[q] = nullptr;
  return ok;
}

int vgroup_barrier(cmb_ctx* c) {
  CMB_CUDA(cudaStreamSynchronize(c->stream));
  if (!vg_barrier(c->vgroup)) return vg_fail("barrier");
  return CMB_OK;
}

// All-to-all of int32 lists: this rank offers d_send, cut at send_off (P+1 entries) into the pieces destined to each
// rank, and receives the piece every rank q destined to it into d_recv + recv_off[q].
int vgroup_alltoallv_i32(cmb_ctx* c, const int32_t* d_send, const int64_t* send_off, int32_t* d_recv, const int64_t* recv_off) {
  VGroup* g = c->vgroup;
  CMB_CUDA(cudaStreamSynchronize(c->stream));
  g->slot[c->rank] = d_send;
  g->offsets[c->rank].assign(send_off, send_off + g->P + 1);
  if (!vg_barrier(g)) return vg_fail("alltoallv");
  int rc = CMB_OK;
  for (int q = 0; q < g->P && rc == CMB_OK; ++q) {
    if (q == c->rank) continue;
    const int64_t a = g->offsets[q][c->rank], b = g->offsets[q][c->rank + 1];
    if (b - a != recv_off[q + 1] - recv_off[q]) {
      set_error("virtual alltoallv: rank %d sends %lld entries to rank %d, which expects %lld", q, (long long)(b - a), c->rank,
                (long long)(recv_off[q + 1] - recv_off[q]));
      rc = CMB_ERR_INVALID;
    } else if (b > a && (cudaMemcpyAsync(d_recv + recv_off[q], static_cast<const int32_t*>(g->slot[q]) + a,
                                         sizeof(int32_t) * size_t(b - a), cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess ||
                         cudaStreamSynchronize(c->stream) != cudaSuccess)) {
      set_error("virtual alltoallv: %s", cudaGetErrorString(cudaGetLastError()));
      rc = CMB_ERR_CUDA;
    }
  }
  if (!vg_barrier(g)) return vg_fail("alltoallv");
  return rc;
}

void vgroup_detach(VGroup* g) {
  std::lock_guard<std::mutex> lk(g->mu);
  if (g->attached > 0) --g->attached;
}

int vgroup_attach(VGroup* g, cmb_ctx* c, int rank) {
  {
    std::lock_guard<std::mutex> lk(g->mu);
    ++g->attached;
  }
  c->vgroup = g;
  c->rank = rank;
  c->nranks = g->P;
  c->num_sms = g->sms[rank];
  {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = g->device;
    CMB_CUDA(cudaMemPoolCreate(&c->mempool, &props));
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(c->mempool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  if (g->green) {
    // the compute stream of this rank runs on its own SM partition
    CUstream s = nullptr;
    if (green_api().GreenCtxStreamCreate(&s, g->gctx[rank], CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
      set_error("cuGreenCtxStreamCreate failed for virtual rank %d", rank);
      return CMB_ERR_CUDA;
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    c->stream = s;
  }
  return CMB_OK;
}

int vgroup_size(const VGroup* g) { return g->P; }
int vgroup_device(const VGroup* g) { return g->device; }

}  // namespace cmb

using namespace cmb;

struct cmb_vgroup {
  VGroup g;
};

extern "C" {

int cmb_vgroup_create(int device, int nranks, cmb_vgroup** out) {
  CMB_REQUIRE(out, "null argument");
  *out = nullptr;
  CMB_REQUIRE(nranks >= 1 && nranks <= kMaxPeers && (nranks & (nranks - 1)) == 0, "nranks must be a power of two <= 16");
  // Effective only when this is the first CUDA call of the process (both variables are read when CUDA initialises):
  // eager module loading, because the lazy loading of a kernel on its first launch synchronises the context and
  // would wait for a peer rank's spinning kernel; enough hardware queues for the streams of all ranks.
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("no CUDA device is visible: libcmpt_b200 has no CPU fallback");
    return CMB_ERR_NO_DEVICE;
  }
  CMB_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  CMB_CUDA(cudaSetDevice(device));
  CMB_CUDA(cudaFree(nullptr));  // make sure the primary context exists before green contexts are carved out of it
  if (nranks > 1) {
    typedef CUresult (*GetModeFn)(CUmoduleLoadingMode*);
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUmoduleLoadingMode mode = CU_MODULE_EAGER_LOADING;
    if (cudaGetDriverEntryPoint("cuModuleGetLoadingMode", &fp, cudaEnableDefault, &q) == cudaSuccess && fp &&
        reinterpret_cast<GetModeFn>(fp)(&mode) == CUDA_SUCCESS && mode != CU_MODULE_EAGER_LOADING) {
      set_error("virtual ranks need CUDA_MODULE_LOADING=EAGER in the environment before CUDA initialises: with lazy "
                "loading the first launch of a kernel synchronises the context while a peer rank's kernel spins");
      return CMB_ERR_UNSUPPORTED;
    }
    cudaGetLastError();
  }
  cudaDeviceProp prop;
  CMB_CUDA(cudaGetDeviceProperties(&prop, device));
  cmb_vgroup* vg = new (std::nothrow) cmb_vgroup();
  if (!vg) return CMB_ERR_NOMEM;
  vg->g.device = device;
  vg->g.P = nranks;
  vg->g.slot.assign(nranks, nullptr);
  vg->g.offsets.resize(nranks);
  if (const char* t = getenv("CMPT_B200_VGROUP_TIMEOUT_S")) vg->g.timeout_s = std::max(1.0, atof(t));
  make_green_contexts(&vg->g, prop.multiProcessorCount);
  *out = vg;
  return CMB_OK;
}

int cmb_vgroup_destroy(cmb_vgroup* vg) {
  if (!vg) return CMB_OK;
  {
    std::lock_guard<std::mutex> lk(vg->g.mu);
    if (vg->g.attached > 0) {  // the contexts keep a pointer to the group and streams in its green contexts
      set_error("cmb_vgroup_destroy: %d context(s) of the group still exist; destroy them first", vg->g.attached);
      return CMB_ERR_INVALID;
    }
  }
  if (vg->g.green)
    for (auto c : vg->g.gctx) green_api().GreenCtxDestroy(c);
  delete vg;
  return CMB_OK;
}

int cmb_vgroup_info(const cmb_vgroup* vg, int* nranks, int* green_contexts, int* sms_per_rank) {
  CMB_REQUIRE(vg, "null argument");
  if (nranks) *nranks = vg->g.P;
  if (green_contexts) *green_contexts = vg->g.green ? 1 : 0;
  if (sms_per_rank) *sms_per_rank = vg->g.sms.empty() ? 0 : vg->g.sms[0];
  return CMB_OK;
}

}  // extern "C"

namespace cmb {
VGroup* vgroup_of(cmb_vgroup* vg) { return vg ? &vg->g : nullptr; }
}  // namespace cmb
