// halo.cuh — row partition + halo exchange interfaces (halo.cu).
#pragma once
#include <algorithm>
#include <vector>

#include "kernels.cuh"

namespace cmb {

int64_t partition_begin(int64_t n, int P, int q);
// host-only: distinct remote columns (sorted), per-owner counts, and the remapped column array
int plan_halo(int64_t n, int P, int rank, int64_t nnz, const int32_t* col, int32_t* col_local,
              std::vector<int32_t>& halo_cols, std::vector<int64_t>& per_owner);

// the same plan computed on the device: d_col (nnz global column indices in HBM) is remapped in place
int plan_halo_device(cmb_ctx* ctx, int64_t n, int P, int rank, int64_t nnz, int32_t* d_col, std::vector<int32_t>& halo_cols,
                     std::vector<int64_t>& per_owner);

struct HaloExchange {
  int P = 1, rank = 0, es = 1;
  int64_t nsend = 0, nrecv = 0;
  std::vector<int64_t> send_off, recv_off;
  int32_t* d_send_idx = nullptr;
  double* d_sendbuf = nullptr;
  double* d_halo = nullptr;
  // peer-memory push path (CUDA IPC over NVLink); NCCL send/recv is the fallback when the mapping is unavailable
  bool p2p = false;
  cmb_ctx* p2p_ctx = nullptr;
  cmb_ctx* own_ctx = nullptr;  // the context setup() ran on
  void* p2p_base = nullptr;  // own allocation: [kMaxPeers flags, padded to 256 B][2 x nrecv*es doubles]
  void* p2p_mapped[kMaxPeers] = {};
  HaloPush push;
  HaloPull pull;
  ~HaloExchange();
  int setup(cmb_ctx* ctx, int64_t n_global, int es, const std::vector<int32_t>& halo_cols,
            const std::vector<int64_t>& per_owner);
  int exchange(cmb_ctx* ctx, const double* w, const int* halt);
  // producer arguments for a consuming kernel launched with `grid` CTAs of 256 threads
  HaloPush fused_push(int grid) const {
    HaloPush h = push;
    const int64_t want = (nsend + 255) / 256;
    h.npush = int(std::max<int64_t>(1, std::min<int64_t>(want, grid)));
    return h;
  }
  // what the consuming kernel needs: which buffer to gather from and which flags to wait on
  HaloPull pull_args() const {
    if (p2p) return pull;
    HaloPull h;
    h.base = d_halo;
    return h;
  }

 private:
  int setup_p2p(cmb_ctx* ctx, const std::vector<double>& cnt);
};

}  // namespace cmb
