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
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "cmpt_b200_debug.h"
#include "device_utils.cuh"
#include "op.cuh"

namespace cmb {

constexpr int kMaxFull = 6;        // rank-rank bonds (p <= 7)

// The local bonds are a nearest-neighbour stencil on the bits of the state index: element s gathers from
// s ^ (3 << b).  One pass of the kernel stages TILES of 2^tb elements (64 KB) in shared memory and serves from a tile
// every bond whose two bits lie inside it.  A tile is the set of indices that share all bits outside
//     {0 .. a-1}  (contiguous run of 256 bytes: keeps global accesses coalesced)   and   {h .. h+tb-a-1}  (the window),
// so the first pass (a = tb, contiguous tiles) covers bonds 0 .. tb-2 and every further pass moves the window up.
// The passes' contributions add up in v; the diagonal term splits over the passes the same way.
// Data movement: persistent CTAs, one producer thread feeding a 3-stage ring of tiles through TMA (a bulk copy for
// the contiguous tiles of the first pass, a 3-D tensor box {run, 1, window} for the others), 16 consumer warps.
struct HeisPass {
  int tb;     // tile bits
  int a;      // contiguous low bits of the tile
  int h;      // global position of tile bit a (start of the window)
  int k0;     // this pass serves the bonds joining tile bits (k, k+1), k0 <= k < tb-1 ...
  int wrapb;  // ... and, when set, the periodic bond joining tile bits tb-1 and 0 (single rank only)
  int main;   // the contiguous pass, run last: writes u, adds the remote bonds and the shift, computes the alpha dot
  int rmw;    // adds to the v of the earlier passes (the first pass of an apply just writes v)
  // Sibling passes (XB = 3): besides the bonds inside its tile a CTA serves the three bonds that involve the three index
  // bits directly above the tile — (top tile bit, x0), (x0, x1), (x1, x2).  Their partner elements lie in the seven
  // "sibling" tiles that differ in those bits; the work order keeps the eight siblings on neighbouring CTAs at the
  // same time, so the partner elements are read straight from global memory and hit in L2 (no second DRAM read).
  int xmask;  // which of these three bonds this pass serves (bit e = bond e)
  int xwrap;  // ... and the periodic bond joining x2 (the top index bit) with tile bit 0 (single rank only)
};

constexpr int kHeisConsumers = 512;  // 16 warps (128 registers each); thread 0 also feeds the ring
constexpr int kHeisThreads = kHeisConsumers;
constexpr int kHeisStages = 3;
constexpr int kHeisTileBytes = 65536;
constexpr int kHeisSmem = kHeisStages * kHeisTileBytes + 2 * kHeisStages * 8;

struct HeisArgs {
  int Ll;             // local bits
  int nb;             // total number of bonds (for the diagonal)
  int aligned_uniform;  // rank-rank bonds whose two rank bits agree
  int n_full;           // rank-rank bonds whose rank bits differ: full partner slabs
  const double* full[kMaxFull];
  int has_straddle, rb0;  // bond (Ll-1, Ll)
  const double* half;
  int has_wrap, rt;  // bond (L-1, 0) across ranks
  const double* wrap;
  double J;
  // peer-memory exchange: partners copy their slabs into this rank's receive buffer and then set flag[k] = seq;
  // the contiguous pass waits for the flags named in flag_mask before it touches the slabs
  const unsigned long long* flag;
  unsigned long long seq;
  unsigned flag_mask;
  int* error;
  long long timeout;  // spin bound of the flag wait in SM clocks (cmb_ctx::spin_timeout)
};

// Two neighbouring tile elements (t0 even, t0 + 1): the unit of work of one thread.  For every bond that does not
// involve tile bit 0 the two share the aligned/anti-aligned decision and their partners are neighbours too, so one
// test and one 16-byte shared-memory load serve both.
template <bool CPLX>
struct HeisPair {
  double r0, i0, r1, i1;
  __device__ __forceinline__ static HeisPair zero() { return HeisPair{0.0, 0.0, 0.0, 0.0}; }
  __device__ __forceinline__ static HeisPair load(const double* base, long long t0) {
    HeisPair p;
    if constexpr (CPLX) {
      const double2 e0 = reinterpret_cast<const double2*>(base)[t0];
      const double2 e1 = reinterpret_cast<const double2*>(base)[t0 + 1];
      p.r0 = e0.x, p.i0 = e0.y, p.r1 = e1.x, p.i1 = e1.y;
    } else {
      const double2 e = *reinterpret_cast<const double2*>(base + t0);
      p.r0 = e.x, p.r1 = e.y, p.i0 = 0.0, p.i1 = 0.0;
    }
    return p;
  }
  __device__ __forceinline__ void store(double* base, long long t0) const {
    if constexpr (CPLX) {
      reinterpret_cast<double2*>(base)[t0] = make_double2(r0, i0);
      reinterpret_cast<double2*>(base)[t0 + 1] = make_double2(r1, i1);
    } else {
      *reinterpret_cast<double2*>(base + t0) = make_double2(r0, r1);
    }
  }
  // streaming variants (L2 evict-first): data that is touched once per pass
  __device__ __forceinline__ static HeisPair load_cs(const double* base, long long t0) {
    HeisPair p;
    if constexpr (CPLX) {
      const double2 e0 = __ldcs(reinterpret_cast<const double2*>(base) + t0);
      const double2 e1 = __ldcs(reinterpret_cast<const double2*>(base) + t0 + 1);
      p.r0 = e0.x, p.i0 = e0.y, p.r1 = e1.x, p.i1 = e1.y;
    } else {
      const double2 e = __ldcs(reinterpret_cast<const double2*>(base + t0));
      p.r0 = e.x, p.r1 = e.y, p.i0 = 0.0, p.i1 = 0.0;
    }
    return p;
  }
  __device__ __forceinline__ void store_cs(double* base, long long t0) const {
    if constexpr (CPLX) {
      __stcs(reinterpret_cast<double2*>(base) + t0, make_double2(r0, i0));
      __stcs(reinterpret_cast<double2*>(base) + t0 + 1, make_double2(r1, i1));
    } else {
      __stcs(reinterpret_cast<double2*>(base + t0), make_double2(r0, r1));
    }
  }
  // this += f * o
  __device__ __forceinline__ void axpy(double f, const HeisPair& o) {
    r0 = fma(f, o.r0, r0), r1 = fma(f, o.r1, r1);
    if constexpr (CPLX) i0 = fma(f, o.i0, i0), i1 = fma(f, o.i1, i1);
  }
  // element 0 += f * o.element 1   /   element 1 += f * o.element 0
  __device__ __forceinline__ void axpy_cross(double f, const HeisPair& o, bool into0) {
    if (into0) {
      r0 = fma(f, o.r1, r0);
      if constexpr (CPLX) i0 = fma(f, o.i1, i0);
    } else {
      r1 = fma(f, o.r0, r1);
      if constexpr (CPLX) i1 = fma(f, o.i0, i1);
    }
  }
  __device__ __forceinline__ void add(const HeisPair& o) {
    r0 += o.r0, r1 += o.r1;
    if constexpr (CPLX) i0 += o.i0, i1 += o.i1;
  }
  __device__ __forceinline__ void add_if(const HeisPair& o, bool c) {
    r0 += c ? o.r0 : 0.0, r1 += c ? o.r1 : 0.0;
    if constexpr (CPLX) i0 += c ? o.i0 : 0.0, i1 += c ? o.i1 : 0.0;
  }
  // element 0 += o.element 1   /   element 1 += o.element 0
  __device__ __forceinline__ void add_cross(const HeisPair& o, bool into0) {
    if (into0) {
      r0 += o.r1;
      if constexpr (CPLX) i0 += o.i1;
    } else {
      r1 += o.r0;
      if constexpr (CPLX) i1 += o.i0;
    }
  }
};

// MAIN / RMW: role of the pass (compile time, so that the unrolled pair loop carries no uniform branches for them)
template <bool CPLX, bool MAIN, bool RMW, int XB>
__global__ void __launch_bounds__(kHeisThreads, 1)
heis_apply_kernel(HeisArgs a, HeisPass ps, const __grid_constant__ CUtensorMap tm, const double* __restrict__ w,
                  double* __restrict__ ucol, double* __restrict__ v, double shr, double shi, StepScalars sc,
                  double* partial, unsigned* ticket) {
  using Pair = HeisPair<CPLX>;
  constexpr int ES = CPLX ? 2 : 1;
  constexpr int PPT = CPLX ? 4 : 8;  // pairs per consumer thread in a full tile
  extern __shared__ __align__(128) unsigned char heis_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(heis_smem + kHeisStages * kHeisTileBytes);
  uint64_t* empty = full + kHeisStages;
  grid_dependency_wait();  // w, v of the earlier passes, the step scalars and the halt flag come from earlier kernels
  grid_launch_dependents();
  double inv;
  if (!step_prologue(sc, inv)) return;  // idempotent: every pass of one apply takes the same decision
  const int tb = ps.tb;
  const int npairs = 1 << (tb - 1);
  const int nmid = ps.h - ps.a;  // index bits between the contiguous run and the window
  const long long ntiles = 1ll << (a.Ll - tb);
  if (MAIN && a.flag != nullptr && threadIdx.x < 32) {
    if ((a.flag_mask >> threadIdx.x) & 1u) {
      const long long t0 = clock64();
      while (ld_acquire_sys_u64(a.flag + threadIdx.x) < a.seq) {
        if (clock64() - t0 > a.timeout) {  // bounded: a missing partner raises the error flag and halts the chain
          if (blockIdx.x == 0)
            printf("[cmpt_b200] Heisenberg slab wait timed out: bond %d, exchange %llu expected, flag holds %llu\n",
                   int(threadIdx.x), a.seq, ld_acquire_sys_u64(a.flag + threadIdx.x));
          *a.error = 1;
          *sc.halt = 1;
          break;
        }
      }
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kHeisStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kHeisConsumers / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  double d0 = 0.0, d1 = 0.0;
  // ring feeder (thread 0): one TMA per tile into stage `s`, signalled on full[s]
  const uint32_t tile_bytes = uint32_t(sizeof(double) * ES) << tb;
  const bool contiguous = (ps.a == tb);
  // linear work index q -> tile.  In a sibling pass the three low bits of q select among the eight sibling tiles, i.e.
  // they must land on the three index bits directly above the tile: for a contiguous tile those are the low bits of
  // the tile number anyway, for a window tile they are the low bits of its "high" part.
  auto tile_of = [&](long long q) -> long long {
    if (!XB || contiguous) return q;
    const long long rest = q >> XB;
    const long long mid = rest & ((1ll << nmid) - 1), high = rest >> nmid;
    return mid | (((high << XB) | (long long)(q & ((1 << XB) - 1))) << nmid);
  };
  // sibling passes: w is read twice (own tile through TMA, partner elements through L2) while v and u stream by once,
  // so w's lines are kept in L2 (evict-last) and the others marked evict-first
  const uint64_t keep = XB ? l2_policy_evict_last() : 0ull;
  auto feed = [&](long long q, int s) {
    const long long tile = tile_of(q);
    mbar_arrive_expect_tx(&full[s], tile_bytes);
    char* dst = reinterpret_cast<char*>(heis_smem) + size_t(s) * kHeisTileBytes;
    if (contiguous) {
      if (XB) bulk_load_1d_hint(dst, w + (size_t(tile) << tb) * ES, tile_bytes, &full[s], keep);
      else bulk_load_1d(dst, w + (size_t(tile) << tb) * ES, tile_bytes, &full[s]);
    } else {
      // the window goes as boxes of at most 256 window positions (the limit of a TMA box dimension)
      const int wpos = 1 << (tb - ps.a), nbox = wpos > 256 ? wpos / 256 : 1;
      const uint32_t box_bytes = tile_bytes / uint32_t(nbox);
      for (int b = 0; b < nbox; ++b) {
        const int c1 = int(tile & ((1ll << nmid) - 1)), c2 = int((tile >> nmid) << (tb - ps.a)) + b * 256;
        if (XB) tma_load_3d_hint(dst + size_t(b) * box_bytes, &tm, 0, c1, c2, &full[s], keep);
        else tma_load_3d(dst + size_t(b) * box_bytes, &tm, 0, c1, c2, &full[s]);
      }
    }
  };
  if (threadIdx.x == 0) {
    if (!contiguous) prefetch_tmap(&tm);
    for (int s = 0; s < kHeisStages; ++s) {
      const long long q = blockIdx.x + (long long)s * gridDim.x;
      if (q < ntiles) feed(q, s);
    }
  }
  {
    const int lowmask = (1 << ps.a) - 1;
    const long long top = 1ll << (a.Ll - 1);
    const double hJ = 0.5 * a.J, qJ = 0.25 * a.J;
    const int kfirst = ps.k0 > 1 ? ps.k0 : 1;
    const int kmask = ((1 << (tb - 1)) - 1) & ~((1 << kfirst) - 1);  // gray-code bits of the bonds k >= kfirst
    const bool bond0 = (ps.k0 == 0 && tb >= 2);
    const int x0b = XB ? (ps.xmask & 1) : 0, x1b = XB ? ((ps.xmask >> 1) & 1) : 0, x2b = XB ? ((ps.xmask >> 2) & 1) : 0;
    const int xwrap = XB ? ps.xwrap : 0;
    const int nlocal = (tb - 1 - ps.k0) + ps.wrapb + x0b + x1b + x2b + xwrap;
    const int xb0 = ps.h + tb - ps.a;  // global position of sibling bit 0
    // bonds whose diagonal term this pass accounts for (the remote bonds ride with the first pass)
    const bool remote = (a.n_full | a.has_straddle | a.has_wrap) != 0;
    const int nb_pass = nlocal + (MAIN ? a.aligned_uniform + a.n_full + a.has_straddle + a.has_wrap : 0);
    // pair g = threadIdx.x + 512 j  ->  tile element t0 = 2 g = tl + (j << 10); the tile -> global index map is a
    // bit deposit, so it splits into a per-thread part (hoisted) and a per-iteration part
    const int tl = 2 * threadIdx.x;  // < 1024
    const int gm = (tl ^ (tl >> 1)) & kmask & 0x1fe;  // bonds 1..8: anti-aligned and served by this pass
    const int anti_tid = __popc(gm);
    const int b9 = (tl >> 9) & 1, b1 = (tl >> 1) & 1;
    const long long s_tid = (long long)(tl & lowmask) | ((long long)(tl >> ps.a) << ps.h);
    auto s_iter = [&](int j) -> long long {
      const int tj = j << 10;
      return (long long)(tj & lowmask) | ((long long)(tj >> ps.a) << ps.h);
    };
    int st = 0, it = 0;
    uint32_t ph = 0;
    for (long long q = blockIdx.x; q < ntiles; q += gridDim.x, ++it) {
      const long long tile = tile_of(q);
      const long long base =
          (((tile & ((1ll << nmid) - 1)) << ps.a) | ((tile >> nmid) << (ps.h + tb - ps.a))) + s_tid;
      // Before waiting for the tile: what this tile's result is added to (the earlier passes' v) and the sibling bonds'
      // partner elements, both from global memory.  yold = v_old + (J/2)/beta * (sum of the sibling partners).
      const unsigned crank = XB ? unsigned(q & ((1 << XB) - 1)) : 0u;
      const bool x1anti = x1b && (((crank ^ (crank >> 1)) & 1u) != 0u);         // sibling bits 0, 1 differ: uniform per tile
      const bool x2anti = x2b && ((((crank >> 1) ^ (crank >> 2)) & 1u) != 0u);  // sibling bits 1, 2 differ
      Pair yold[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        yold[j] = Pair::zero();
        if (RMW && int(threadIdx.x) + kHeisConsumers * j < npairs)
          yold[j] = XB ? Pair::load_cs(v, base + s_iter(j)) : Pair::load(v, base + s_iter(j));
      }
      if (XB) {  // tiles are full here (tb = TB).  One loop per bond: its PPT loads are independent and go out together
        const double f = hJ * inv;
        if (x1anti) {
#pragma unroll
          for (int j = 0; j < PPT; ++j) yold[j].axpy(f, Pair::load(w, (base + s_iter(j)) ^ (3ll << xb0)));
        }
        if (x2anti) {
#pragma unroll
          for (int j = 0; j < PPT; ++j) yold[j].axpy(f, Pair::load(w, (base + s_iter(j)) ^ (3ll << (xb0 + 1))));
        }
        if (x0b) {
          const int c0 = int(crank & 1u);
#pragma unroll
          for (int j = 0; j < PPT; ++j)
            if ((((tl | (j << 10)) >> (tb - 1)) & 1) != c0)
              yold[j].axpy(f, Pair::load(w, (base + s_iter(j)) ^ (3ll << (xb0 - 1))));
        }
        if (xwrap) {
#pragma unroll
          for (int j = 0; j < PPT; ++j)
            yold[j].axpy_cross(f, Pair::load(w, (base + s_iter(j)) ^ (1ll << (xb0 + 2))), (crank & 4u) != 0u);
        }
      }
      mbar_wait(&full[st], ph);
      const double* sm = reinterpret_cast<const double*>(heis_smem + size_t(st) * kHeisTileBytes);
      // This thread's pairs are t0(j) = tl | (j << 10): tile bits 0..9 are the thread's, bits 10.. are j (compile
      // time after unrolling).  Bonds k <= 8 therefore have a per-thread aligned/anti-aligned pattern (gm) and one
      // test serves all PPT pairs; bonds 10 and 11 are decided by j alone.  Loads may touch stage bytes outside a
      // small tile (tb < 13) for pairs that do not exist; those results are never stored.
      Pair acc[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) acc[j] = Pair::zero();
#pragma unroll
      for (int k = 1; k <= 8; ++k) {
        if (gm & (1 << k)) {
          const int po = tl ^ (3 << k);
#pragma unroll
          for (int j = 0; j < PPT; ++j) acc[j].add(Pair::load(sm, po | (j << 10)));
        }
      }
      if (kmask & (1 << 9)) {  // tile bits 9 (thread) and 10 (j bit 0)
        const int po = tl ^ (1 << 9);
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (b9 != (j & 1)) acc[j].add(Pair::load(sm, po | ((j ^ 1) << 10)));
      }
      if (kmask & (1 << 10)) {  // tile bits 10, 11 = j bits 0, 1
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (((j ^ (j >> 1)) & 1) != 0) acc[j].add(Pair::load(sm, tl | ((j ^ 3) << 10)));
      }
      if (PPT > 4 && (kmask & (1 << 11))) {  // tile bits 11, 12 = j bits 1, 2
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (((j ^ (j >> 1)) & 2) != 0) acc[j].add(Pair::load(sm, tl | ((j ^ 6) << 10)));
      }
      if (bond0) {  // tile bits 0 and 1: exactly one of the two elements of a pair is anti-aligned
        const int po = tl ^ 2;
#pragma unroll
        for (int j = 0; j < PPT; ++j) acc[j].add_cross(Pair::load(sm, po | (j << 10)), b1 != 0);
      }
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int g = int(threadIdx.x) + kHeisConsumers * j;
        if (g < npairs) {
          const int t0 = tl | (j << 10);
          const long long s0 = base + s_iter(j);
          const Pair x = Pair::load(sm, t0);
          // anti-aligned bonds: thread part + bond 9 + the j-only bonds 10, 11 (+ bond 0, per element)
          int anti0 = anti_tid + ((kmask >> 9) & (b9 ^ (j & 1))) + __popc(((j ^ (j >> 1)) << 10) & kmask);
          int anti1 = anti0;
          if (bond0) {
            anti0 += b1;
            anti1 += 1 - b1;
          }
          if (ps.wrapb) {  // tile bits tb-1 and 0 (single rank, periodic chain)
            const int bt = (t0 >> (tb - 1)) & 1;
            acc[j].add_cross(Pair::load(sm, t0 ^ (1 << (tb - 1))), bt != 0);
            anti0 += bt;
            anti1 += 1 - bt;
          }
          if (XB) {  // sibling bonds: the partner elements are in yold already, here only the diagonal counts
            const int bt = (t0 >> (tb - 1)) & 1;
            const int ax = ((x0b && bt != int(crank & 1u)) ? 1 : 0) + (x1anti ? 1 : 0) + (x2anti ? 1 : 0);
            anti0 += ax;
            anti1 += ax;
            if (xwrap) {  // periodic bond: sibling bit 2 (the top index bit) and index bit 0
              const int c2 = int((crank >> 2) & 1u);
              anti0 += c2;
              anti1 += 1 - c2;
            }
          }
          if (MAIN && remote) {
            anti0 += a.n_full;
            anti1 += a.n_full;
#pragma unroll
            for (int k = 0; k < kMaxFull; ++k)
              if (k < a.n_full) acc[j].add(Pair::load(a.full[k], s0));
            if (a.has_straddle) {
              if (int((s0 >> (a.Ll - 1)) & 1) != a.rb0) {
                acc[j].add(Pair::load(a.half, s0 & (top - 1)));
                ++anti0;
                ++anti1;
              }
            }
            if (a.has_wrap) {  // partner element (s >> 1) of the packed half slab, for the element whose bit 0 != rt
              const double* q = a.wrap + (s0 >> 1) * ES;
              if (a.rt) {
                acc[j].r0 += q[0];
                if constexpr (CPLX) acc[j].i0 += q[1];
                ++anti0;
              } else {
                acc[j].r1 += q[0];
                if constexpr (CPLX) acc[j].i1 += q[1];
                ++anti1;
              }
            }
          }
          const double dg0 = qJ * double(nb_pass - 2 * anti0), dg1 = qJ * double(nb_pass - 2 * anti1);
          Pair u, y;
          u.r0 = x.r0 * inv, u.r1 = x.r1 * inv, u.i0 = x.i0 * inv, u.i1 = x.i1 * inv;
          y.r0 = (dg0 * x.r0 + hJ * acc[j].r0) * inv;
          y.r1 = (dg1 * x.r1 + hJ * acc[j].r1) * inv;
          y.i0 = y.i1 = 0.0;
          if constexpr (CPLX) {
            y.i0 = (dg0 * x.i0 + hJ * acc[j].i0) * inv;
            y.i1 = (dg1 * x.i1 + hJ * acc[j].i1) * inv;
          }
          if (MAIN) {
            y.r0 += shr * u.r0, y.r1 += shr * u.r1;
            if constexpr (CPLX) {
              y.r0 -= shi * u.i0, y.r1 -= shi * u.i1;
              y.i0 += shr * u.i0 + shi * u.r0, y.i1 += shr * u.i1 + shi * u.r1;
            }
            if (XB) u.store_cs(ucol, s0);
            else u.store(ucol, s0);
          }
          if (RMW || XB) y.add(yold[j]);
          if (XB) y.store_cs(v, s0);
          else y.store(v, s0);
          if (MAIN) {
            d0 = fma(u.r0, y.r0, d0);
            d0 = fma(u.r1, y.r1, d0);
            if constexpr (CPLX) {
              d0 += u.i0 * y.i0 + u.i1 * y.i1;
              d1 += u.r0 * y.i0 - u.i0 * y.r0 + u.r1 * y.i1 - u.i1 * y.r1;
            }
          }
        }
      }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[st]);
      if (threadIdx.x == 0) {  // refill this stage with the tile kHeisStages ahead once all 16 warps released it
        const long long nq = q + (long long)kHeisStages * gridDim.x;
        if (nq < ntiles) {
          mbar_wait(&empty[st], ph);
          feed(nq, st);
        }
      }
      if (++st == kHeisStages) {
        st = 0;
        ph ^= 1u;
      }
    }
  }
  if (MAIN) grid_sum_finalize<CPLX ? 2 : 1>(d0, d1, partial, ticket, sc.alpha_slot);
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
  std::vector<HeisPass> passes;
  const bool trace_slab = getenv("CMPT_B200_TRACE_SLAB") != nullptr;
  bool use_siblings = false;  // passes also serve the three bonds above the tile from the sibling tiles (see HeisPass)
  // peer-memory exchange (CUDA IPC + copy engines): every rank owns [flags | receive buffer 0 | receive buffer 1],
  // mapped into its partners, which copy their slabs into it (cudaMemcpyAsync on side streams, overlapping the window
  // passes) and then publish the exchange number in the flag of that bond.  NCCL send/recv is the fallback.
  static constexpr int kMaxRemote = kMaxFull + 2;
  static constexpr size_t kFlagBytes = 256;
  static constexpr int kSeqSlots = 4096;
  bool p2p = false;
  void* p2p_base = nullptr;
  void* p2p_mapped[kMaxPeers] = {};
  size_t p2p_tot = 0;                           // doubles per receive buffer of this rank
  std::vector<size_t> peer_off, peer_tot;       // per remote bond: where my slab goes in the partner's buffer / its size
  unsigned long long xseq = 0;                  // exchanges enqueued so far (same on every rank)
  // Bonds whose slab travels by copy engine even when the other slabs are stored by the pass that produces w
  // (CMPT_B200_SLAB_CE_SECOND=1, off by default).  A rank with two anti-aligned rank-rank bonds takes in 3 slabs' worth
  // of NVLink stores during one pass — more than the links carry in that time; with this option its second full slab
  // (and the partner's, for symmetry) waits for the end of the pass and travels while the window passes run.  Measured
  // at L = 30 on 8 GPUs: UPDATE_NORM on those ranks drops from 5.3 to 4.4 ms, but their apply then waits as long for the
  // copy (64.7 against 65.4 it/s) — the exchange is bound by what a rank can take in, whichever engine moves it.
  unsigned ce_bonds = 0;
  const double* pushed_w = nullptr;             // slab_push_begin(): the producer of this w pushes exchange pushed_x
  unsigned long long pushed_x = 0;
  unsigned long long* h_seq = nullptr;          // pinned source words of the flag copies
  static constexpr int kMaxSplit = 4;            // a slab travels as up to 4 concurrent copies (one copy engine each)
  int nsplit = 2;
  cudaStream_t xstream[kMaxRemote][kMaxSplit] = {};
  cudaEvent_t ev_ready = nullptr, ev_done[kMaxRemote] = {}, ev_chunk[kMaxRemote][kMaxSplit] = {};
  ~HeisenbergOp() override {
    if (p2p) {
      cudaStreamSynchronize(ctx->stream);
      for (int i = 0; i < kMaxRemote; ++i) {
        for (int c = 0; c < kMaxSplit; ++c) {
          if (xstream[i][c]) {
            cudaStreamSynchronize(xstream[i][c]);
            cudaStreamDestroy(xstream[i][c]);
          }
          if (ev_chunk[i][c]) cudaEventDestroy(ev_chunk[i][c]);
        }
        if (ev_done[i]) cudaEventDestroy(ev_done[i]);
      }
      if (ev_ready) cudaEventDestroy(ev_ready);
      ipc_unshare(ctx, p2p_mapped);
      rank_barrier(ctx);  // collective: nobody frees an exported buffer a peer may still write to
      dfree(ctx, p2p_base);
      hfree(ctx, h_seq);
    }
    if (ctx) {
      pool_free(ctx, d_recv);
      pool_free(ctx, d_pack);
    }
  }

  HeisArgs base_args() const {
    HeisArgs a;
    memset(&a, 0, sizeof(a));
    a.Ll = plan.Ll;
    a.nb = plan.nb;
    a.J = J;
    return a;
  }

  // Tiling of the local bonds (see HeisPass).  64 KB tiles: 2^13 doubles or 2^12 complex numbers; window passes
  // keep runs of 256 bytes and move an 8-bit window (box rows of a TMA tensor map are limited to 256).
  void plan_passes() {
    passes.clear();
    const int Ll = plan.Ll, TB = cplx ? 12 : 13;
    const bool local_wrap = (plan.p == 0 && plan.nb == plan.L);
    // Sibling passes (CMPT_B200_HEIS_SIBLINGS=1, off by default) need three index bits above the 2^TB tile; their window
    // passes use 128-byte runs so that window + sibling bits reach 12 bits: two passes up to 27 local bits instead of
    // three.  Measured on B200 at 27 bits the two passes move 1/3 fewer DRAM bytes but are bound by the latency of the
    // partner loads (2.0 ms per apply against 1.7 ms for the three HBM-bound passes; profiles/r2/heis_two_pass.md), so
    // the default keeps every pass inside its own tiles (256-byte runs).
    const char* sib = getenv("CMPT_B200_HEIS_SIBLINGS");
    use_siblings = (Ll >= TB + 3) && sib && atoi(sib) != 0;
    const int XB = use_siblings ? 3 : 0;
    const int arun = use_siblings ? (cplx ? 3 : 4) : (cplx ? 4 : 5);
    const int wbl = TB - arun;  // window bits inside one tile
    const int wb = wbl + XB;    // index bits a window pass spans above h
    HeisPass m;  // the contiguous pass: bonds 0 .. tb-2 (+ the sibling bonds); runs last because it also consumes the remote slabs
    memset(&m, 0, sizeof(m));
    m.tb = std::min(Ll, TB);
    m.a = m.h = m.tb;
    m.wrapb = (Ll <= TB && local_wrap) ? 1 : 0;
    m.main = 1;
    int next = m.tb - 1;  // first bond not served yet (bond b joins bits b and b+1)
    if (use_siblings) {
      m.xmask = 7;  // bonds tb-1, tb, tb+1
      m.xwrap = (m.tb + XB == Ll && local_wrap) ? 1 : 0;
      next = m.tb + XB - 1;
    }
    while (next <= Ll - 2) {
      HeisPass q;
      memset(&q, 0, sizeof(q));
      q.tb = TB;
      q.a = arun;
      q.h = std::min(next, Ll - wb);  // the last window is pulled down so that it ends at the top bit
      q.k0 = q.a + (next - q.h);      // tile bit of the first bond to serve (may lie beyond the tile's own window)
      if (q.k0 > q.tb - 1) q.k0 = q.tb - 1;
      if (use_siblings) {
        // sibling bond e joins index bits h + wbl - 1 + e and h + wbl + e: served when it is not below `next`
        for (int e = 0; e < XB; ++e)
          if (q.h + wbl - 1 + e >= next) q.xmask |= 1 << e;
        q.xwrap = (q.h + wb == Ll && local_wrap) ? 1 : 0;
      } else {
        q.wrapb = (q.h + wb == Ll && local_wrap) ? 1 : 0;
      }
      q.rmw = passes.empty() ? 0 : 1;
      next = q.h + wb - 1;
      passes.push_back(q);
    }
    m.rmw = passes.empty() ? 0 : 1;
    passes.push_back(m);
  }

  // 3-D view of a local vector for the tiles of a window pass: {run, index bits between run and window, the rest}
  int window_map(const HeisPass& ps, const double* w, CUtensorMap* tm) const {
    const int es = cplx ? 2 : 1;
    cuuint64_t gdim[3] = {cuuint64_t(es) << ps.a, cuuint64_t(1) << (ps.h - ps.a), cuuint64_t(1) << (plan.Ll - ps.h)};
    cuuint64_t gstr[2] = {(cuuint64_t(8 * es) << ps.a), (cuuint64_t(8 * es) << ps.h)};
    cuuint32_t box[3] = {cuuint32_t(es) << ps.a, 1u, std::min<cuuint32_t>(1u << (ps.tb - ps.a), 256u)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = get_encode_tiled()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(w), gdim, gstr, box,
                                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed (%d) for the Heisenberg window a=%d h=%d Ll=%d", int(cr), ps.a, ps.h,
                plan.Ll);
      return CMB_ERR_CUDA;
    }
    return CMB_OK;
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
    CMB_TRY(pool_alloc(ctx, &d_pack, sizeof(double) * std::max<size_t>(half, 2)));
    CMB_TRY(setup_p2p(tot));
    if (!p2p) CMB_TRY(pool_alloc(ctx, &d_recv, sizeof(double) * std::max<size_t>(tot, 2)));
    return CMB_OK;
  }

  static size_t plan_offsets(const HeisPlan& pl, size_t slab, std::vector<size_t>* off) {
    size_t tot = 0;
    for (auto& r : pl.remote) {
      if (off) off->push_back(tot);
      if (r.needed) tot += (r.kind == 3) ? slab : slab / 2;
    }
    return tot;
  }

  // Collective (every rank of the context creates the operator).  Failure of any step leaves p2p == false.
  int setup_p2p(size_t tot) {
    if (!ctx->mail_ok || getenv("CMPT_B200_NO_P2P_HALO") || int(plan.remote.size()) > kMaxRemote) return CMB_OK;
    const size_t es = cplx ? 2 : 1, slab = size_t(n_local) * es;
    const size_t bytes = kFlagBytes + 2 * sizeof(double) * std::max<size_t>(tot, 2);
    void* base = nullptr;
    if (cudaMalloc(&base, bytes) != cudaSuccess) {
      cudaGetLastError();
      base = nullptr;
    } else {
      cudaMemsetAsync(base, 0, kFlagBytes, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
    }
    void* mapped[kMaxPeers];
    if (!ipc_share(ctx, base, mapped)) {
      dfree(ctx, base);
      cudaGetLastError();
      return CMB_OK;
    }
    bool ok = cudaHostAlloc(reinterpret_cast<void**>(&h_seq), sizeof(unsigned long long) * kSeqSlots,
                            cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming) == cudaSuccess;
    if (const char* e = getenv("CMPT_B200_XCHG_SPLIT")) nsplit = std::max(1, std::min(kMaxSplit, atoi(e)));
    for (size_t k = 0; k < plan.remote.size() && ok; ++k) {
      ok = cudaEventCreateWithFlags(&ev_done[k], cudaEventDisableTiming) == cudaSuccess;
      for (int c = 0; c < nsplit && ok; ++c)
        ok = cudaStreamCreateWithFlags(&xstream[k][c], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ev_chunk[k][c], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {  // local resource failure after the collective part: cannot fall back on one rank only
      set_error("Heisenberg peer exchange: stream/event/pinned allocation failed");
      return CMB_ERR_CUDA;
    }
    p2p = true;
    p2p_base = base;
    p2p_tot = std::max<size_t>(tot, 2);
    for (int q = 0; q < ctx->nranks; ++q) p2p_mapped[q] = mapped[q];
    peer_off.clear();
    peer_tot.clear();
    for (size_t k = 0; k < plan.remote.size(); ++k) {
      HeisPlan pp;
      pp.build(plan.L, plan.pbc, ctx->nranks, plan.remote[k].partner);
      std::vector<size_t> off;
      const size_t ptot = plan_offsets(pp, slab, &off);
      peer_off.push_back(off[k]);  // both partners enumerate the remote bonds in the same order
      peer_tot.push_back(std::max<size_t>(ptot, 2));
    }
    // which full slabs go by copy engine (see ce_bonds): rank-rank bond b when this rank or its partner across that bond
    // has another anti-aligned rank-rank bond below b.  Both partners evaluate the same predicate.  Virtual ranks have no
    // copy engines to spare (same-device copies run on the SMs the peers spin on).
    ce_bonds = 0;
    if (!ctx->vgroup && getenv("CMPT_B200_SLAB_CE_SECOND")) {
      auto differs = [](int r, int b) { return (((r >> b) ^ (r >> (b + 1))) & 1) != 0; };
      for (size_t k = 0; k < plan.remote.size(); ++k) {
        const HeisRemote& r = plan.remote[k];
        if (r.kind != 3 || !r.needed) continue;
        const int b = int(k) - 1;  // remote[0] is the straddle bond, remote[1 + b] joins rank bits b and b + 1
        for (int lower = 0; lower < b; ++lower)
          if (differs(plan.rank, lower) || differs(r.partner, lower)) ce_bonds |= 1u << k;
      }
    }
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
    if (passes.empty()) plan_passes();
    for (const HeisPass& ps : passes) {
      CUtensorMap tm;
      memset(&tm, 0, sizeof(tm));
      if (ps.a != ps.tb) CMB_TRY(window_map(ps, w, &tm));
      LaunchScope ls(ctx, "heisenberg_mf");
      const long long ntiles = 1ll << (plan.Ll - ps.tb);
      int grid = int(std::min<long long>(ntiles, (long long)ctx->num_sms));
#define CMB_HEIS_LAUNCH(C, M, R)                                                                                     \
  do {                                                                                                               \
    auto kern = use_siblings ? heis_apply_kernel<C, M, R, 3> : heis_apply_kernel<C, M, R, 0>;                        \
    static bool attr[2][64] = {};  /* function attributes are per device */                                         \
    if (!attr[use_siblings ? 1 : 0][ctx->device & 63]) {                                                             \
      CMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeisSmem));                  \
      attr[use_siblings ? 1 : 0][ctx->device & 63] = true;                                                           \
    }                                                                                                                \
    CMB_CUDA(launch_pdl(pdl_wanted(true, ctx->nranks), kern, grid, kHeisThreads, kHeisSmem, ctx->stream, a, ps, tm, w, ucol, v, shr, shi, sc,        \
                        ctx->d_partial, ctx->d_ticket + 1));                                                         \
  } while (0)
      const int role = (ps.main ? 2 : 0) | (ps.rmw ? 1 : 0);
      if (cplx) {
        switch (role) {
          case 3: CMB_HEIS_LAUNCH(true, true, true); break;
          case 2: CMB_HEIS_LAUNCH(true, true, false); break;
          case 1: CMB_HEIS_LAUNCH(true, false, true); break;
          default: CMB_HEIS_LAUNCH(true, false, false); break;
        }
      } else {
        switch (role) {
          case 3: CMB_HEIS_LAUNCH(false, true, true); break;
          case 2: CMB_HEIS_LAUNCH(false, true, false); break;
          case 1: CMB_HEIS_LAUNCH(false, false, true); break;
          default: CMB_HEIS_LAUNCH(false, false, false); break;
        }
      }
#undef CMB_HEIS_LAUNCH
      CMB_CUDA(cudaGetLastError());
    }
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

  // Destinations of this rank's slabs for the next exchange (see SlabPush, kernels.cuh): the kernel that writes w stores
  // them through NVLink while it runs, so the exchange costs no time of its own.  Peer-memory exchange only.
  bool slab_push_begin(const double* w, SlabPush* out) override {
    if (plan.p == 0 || !p2p || getenv("CMPT_B200_NO_FUSED_SLAB")) return false;
    const size_t es = cplx ? 2 : 1;
    const size_t slab = size_t(n_local) * es, half = slab / 2;
    const int rb0 = plan.rank & 1, rt = (plan.rank >> (plan.p - 1)) & 1;
    const unsigned long long x = ++xseq;
    const size_t par = size_t(x & 1ull);
    SlabPush sp;
    sp.es = int(es);
    sp.nd = (long long)slab;
    sp.seq = x;
    sp.ticket = ctx->d_ticket + 3;
    for (size_t k = 0; k < plan.remote.size(); ++k) {
      const HeisRemote& r = plan.remote[k];
      if (!r.needed || (ce_bonds & (1u << k))) continue;
      if (sp.n >= kMaxSlabDst) return false;
      char* pbase = static_cast<char*>(p2p_mapped[r.partner]);
      const int d = sp.n++;
      sp.dst[d] = reinterpret_cast<double*>(pbase + kFlagBytes) + par * peer_tot[k] + peer_off[k];
      sp.flag[d] = reinterpret_cast<unsigned long long*>(pbase) + k;
      if (r.kind == 2) {  // straddle: the half whose top local bit differs from my rank bit 0
        sp.kind[d] = 0;
        sp.lo[d] = (long long)(size_t(1 - rb0) * half);
        sp.hi[d] = sp.lo[d] + (long long)half;
      } else if (r.kind == 3) {  // rank-rank bond: the whole slab
        sp.kind[d] = 0;
        sp.lo[d] = 0;
        sp.hi[d] = (long long)slab;
      } else {  // periodic wrap: the elements whose bit 0 differs from my top rank bit, packed
        sp.kind[d] = 1;
        sp.lo[d] = 1 - rt;
        sp.hi[d] = 0;
      }
    }
    pushed_w = w;
    pushed_x = x;
    *out = sp;
    if (trace_slab) fprintf(stderr, "[slab] rank %d announces exchange %llu of %p to %d partners\n", plan.rank, x, (const void*)w, sp.n);
    return true;
  }
  void slab_push_cancel() override { pushed_w = nullptr; }

  int apply(const double* w, double* ucol, double* v, double shr, double shi, const StepScalars& sc) override {
    HeisArgs a = base_args();
    if (plan.p > 0) {
      const bool fused = p2p && pushed_w == w && pushed_w != nullptr;
      if (p2p && !fused && (ctx->vgroup || getenv("CMPT_B200_SM_SLAB_PUSH"))) {
        // virtual ranks (and on request): the exchange of this apply as SM stores from a stand-alone kernel in stream
        // order — same destinations, flags and parity as the fused push
        SlabPush sp;
        if (slab_push_begin(w, &sp)) CMB_TRY(slab_push(ctx, w, sp, sc.halt));
      }
      if (p2p && pushed_w == w && pushed_w != nullptr) {
        // the kernel that produced w (or the push kernel above) stores the slabs and raises the flags of exchange pushed_x
        pushed_w = nullptr;
        const unsigned mask = needed_mask();
        if (ce_bonds & mask) CMB_TRY(copy_bonds(w, pushed_x, ce_bonds & mask));
        const size_t par = size_t(pushed_x & 1ull);
        const double* recv = reinterpret_cast<const double*>(static_cast<char*>(p2p_base) + kFlagBytes) + par * p2p_tot;
        remote_args(a, recv);
        a.flag = static_cast<const unsigned long long*>(p2p_base);
        a.seq = pushed_x;
        a.flag_mask = mask;
        if (trace_slab) fprintf(stderr, "[slab] rank %d consumes exchange %llu (mask %x)\n", plan.rank, pushed_x, mask);
        a.error = ctx->d_mail_error;
        a.timeout = ctx->spin_timeout;
        CMB_TRY(launch(a, w, ucol, v, shr, shi, sc));
        return after_copies(ce_bonds & mask);
      }
      pushed_w = nullptr;
      const size_t es = cplx ? 2 : 1;
      const size_t slab = size_t(n_local) * es, half = slab / 2;
      // what each partner needs from me: straddle -> my half whose top local bit != my rank bit 0 (the partner's
      // rank bit 0); wrap -> my elements whose bit 0 != my top rank bit (packed)
      const int rb0 = plan.rank & 1, rt = (plan.rank >> (plan.p - 1)) & 1;
      bool has_wrap = false;
      for (auto& r : plan.remote) has_wrap |= (r.kind == 4);
      if (has_wrap) CMB_TRY(pack_wrap(w, 1 - rt, d_pack, sc.halt));
      if (p2p) return apply_p2p(a, w, ucol, v, shr, shi, sc);
      if (ctx->vgroup) {
        set_error("virtual ranks exchange slabs through peer memory only");
        return CMB_ERR_UNSUPPORTED;
      }
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

  // Exchange through peer memory.  The copies are unconditional (a halted chain still moves its stale w: the
  // kernels are no-ops then, and every rank keeps the same exchange count).  Receive buffers alternate with the
  // exchange number: a partner can only start exchange x+2 after it consumed my exchange x+1, which I sent after my
  // contiguous pass of exchange x, so the buffer it overwrites is no longer being read.
  // Copies the slabs of the bonds in `bonds` (bit k = remote bond k) of exchange x into the partners' receive buffers and
  // raises their flags.  Virtual ranks (one device): same-device copies run on SMs, and the SMs a side stream could use
  // are held by the peers' spinning kernels, so there the slabs are copied in stream order on the rank's own stream (no
  // overlap, same buffers, flags and parity).  Real ranks: copy engines on side streams, overlapped with whatever the
  // main stream does next (the window passes).
  int copy_bonds(const double* w, unsigned long long x, unsigned bonds) {
    const size_t es = cplx ? 2 : 1;
    const size_t slab = size_t(n_local) * es, half = slab / 2;
    const int rb0 = plan.rank & 1;
    const size_t par = size_t(x & 1ull);
    const bool inl = ctx->vgroup != nullptr;
    if ((x & 255ull) == 0) {  // keeps the pinned flag words of exchanges still in flight from being reused
      if (inl) CMB_CUDA(cudaStreamSynchronize(ctx->stream));
      else
        for (size_t k = 0; k < plan.remote.size(); ++k) CMB_CUDA(cudaStreamSynchronize(xstream[k][0]));
    }
    if (!inl) CMB_CUDA(cudaEventRecord(ev_ready, ctx->stream));  // w (and the packed wrap half) are final
    for (size_t k = 0; k < plan.remote.size(); ++k) {
      const HeisRemote& r = plan.remote[k];
      if (!(bonds & (1u << k))) continue;
      const double* src = w;
      size_t count = slab;
      if (r.kind == 2) {
        src = w + size_t(1 - rb0) * half;
        count = half;
      } else if (r.kind == 4) {
        src = d_pack;
        count = half;
      }
      char* pbase = static_cast<char*>(p2p_mapped[r.partner]);
      double* dst = reinterpret_cast<double*>(pbase + kFlagBytes) + par * peer_tot[k] + peer_off[k];
      unsigned long long* word = h_seq + ((x * kMaxRemote + k) % kSeqSlots);
      *word = x;
      // the slab goes as nsplit concurrent copies; the flag follows on stream 0 once all of them are done
      const int ns = (!inl && count >= (size_t(1) << 16)) ? nsplit : 1;
      const size_t chunk = ((count + ns - 1) / ns + 1) & ~size_t(1);
      for (int c = 0; c < ns; ++c) {
        const size_t b = std::min(count, size_t(c) * chunk), e = std::min(count, b + chunk);
        cudaStream_t cs = inl ? ctx->stream : xstream[k][c];
        if (!inl) CMB_CUDA(cudaStreamWaitEvent(cs, ev_ready, 0));
        if (e > b) CMB_CUDA(cudaMemcpyAsync(dst + b, src + b, sizeof(double) * (e - b), cudaMemcpyDefault, cs));
        if (!inl && c > 0) CMB_CUDA(cudaEventRecord(ev_chunk[k][c], cs));
      }
      for (int c = 1; c < ns; ++c) CMB_CUDA(cudaStreamWaitEvent(xstream[k][0], ev_chunk[k][c], 0));
      CMB_CUDA(cudaMemcpyAsync(reinterpret_cast<unsigned long long*>(pbase) + k, word, sizeof(unsigned long long),
                               cudaMemcpyDefault, inl ? ctx->stream : xstream[k][0]));
      if (!inl) CMB_CUDA(cudaEventRecord(ev_done[k], xstream[k][0]));
    }
    return CMB_OK;
  }
  // whoever overwrites w or the pack buffer next must come after the copies that read them
  int after_copies(unsigned bonds) {
    if (ctx->vgroup) return CMB_OK;
    for (size_t k = 0; k < plan.remote.size(); ++k)
      if (bonds & (1u << k)) CMB_CUDA(cudaStreamWaitEvent(ctx->stream, ev_done[k], 0));
    return CMB_OK;
  }
  unsigned needed_mask() const {
    unsigned mask = 0;
    for (size_t k = 0; k < plan.remote.size(); ++k)
      if (plan.remote[k].needed) mask |= 1u << k;
    return mask;
  }

  int apply_p2p(HeisArgs& a, const double* w, double* ucol, double* v, double shr, double shi, const StepScalars& sc) {
    const unsigned long long x = ++xseq;
    const size_t par = size_t(x & 1ull);
    const unsigned mask = needed_mask();
    CMB_TRY(copy_bonds(w, x, mask));
    const double* recv = reinterpret_cast<const double*>(static_cast<char*>(p2p_base) + kFlagBytes) + par * p2p_tot;
    remote_args(a, recv);
    a.flag = static_cast<const unsigned long long*>(p2p_base);
    a.seq = x;
    a.flag_mask = mask;
    a.error = ctx->d_mail_error;
    a.timeout = ctx->spin_timeout;
    CMB_TRY(launch(a, w, ucol, v, shr, shi, sc));
    return after_copies(mask);
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
  // Virtual ranks share one device: the allocations above synchronise it, so a rank that already launched its first
  // (spinning) apply would stall the ranks still allocating.  Everybody leaves the constructor together.
  if (ctx->vgroup && ctx->nranks > 1) CMB_TRY(rank_barrier(ctx));
  *out = op;
  return CMB_OK;
}

int cmb_heisenberg_plan(int L, int pbc, int nranks, int rank, int* kind, int* partner, int* needed, int* offset_half,
                        int* length_half) {
  if (!kind || !partner || !needed || !offset_half || !length_half) return CMB_ERR_INVALID;
  if (L < 2 || nranks < 1 || (nranks & (nranks - 1)) != 0 || rank < 0 || rank >= nranks) return CMB_ERR_INVALID;
  HeisPlan pl;
  pl.build(L, pbc != 0, nranks, rank);
  if (pl.Ll < 2 || int(pl.remote.size()) > 8) return CMB_ERR_INVALID;
  int off = 0;
  for (size_t k = 0; k < pl.remote.size(); ++k) {
    const HeisRemote& r = pl.remote[k];
    const int len = !r.needed ? 0 : (r.kind == 3 ? 2 : 1);
    kind[k] = r.kind;
    partner[k] = r.partner;
    needed[k] = r.needed ? 1 : 0;
    offset_half[k] = off;
    length_half[k] = len;
    off += len;
  }
  return int(pl.remote.size());
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
