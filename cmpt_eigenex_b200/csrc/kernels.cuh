// kernels.cuh — internal launcher interfaces shared by the translation units of libcmpt_b200.so.
#pragma once
#include "common.cuh"

namespace cmb {

// ---- cgs.cu ------------------------------------------------------------------------------------------
// Fused compute + exchange for the matrix-free Heisenberg operator: the UPDATE_NORM pass that produces the next
// Krylov vector w also stores the parts of w its partner ranks need straight into their receive buffers (NVLink
// stores), tile by tile while it streams the basis, and the CTA that finishes last raises the partners' flags.  The
// operator apply that follows then finds its remote slabs in place instead of waiting for a copy-engine exchange.
constexpr int kMaxSlabDst = 8;
struct SlabPush {
  int n = 0;                              // destinations
  int es = 1;                             // doubles per element
  long long nd = 0;                       // valid doubles of the local vector (the rest of ld is padding)
  int kind[kMaxSlabDst];                  // 0: contiguous range of doubles [lo, hi)   1: every other element, parity lo
  long long lo[kMaxSlabDst], hi[kMaxSlabDst];
  double* dst[kMaxSlabDst];               // where double lo (kind 0) / packed element 0 (kind 1) goes in the peer's buffer
  unsigned long long* flag[kMaxSlabDst];  // the peer's flag word of this bond
  unsigned long long seq = 0;             // exchange number published in the flags
  unsigned* ticket = nullptr;             // device counter of finished CTAs
};

struct CgsPass {
  const double* V = nullptr;  // first column of the chunk
  int64_t ld = 0;             // padded column length in doubles (multiple of 512)
  int64_t col_stride = 0;     // distance between consecutive chunk columns in doubles (ld, or interval*ld)
  int ncols = 0;              // columns in the chunk (scalar columns)
  const double* x = nullptr;  // input vector (nullptr = zero vector)
  double* y = nullptr;        // output vector (modes 1, 2); may alias x
  const double* hin = nullptr;  // coefficients to subtract (modes 1, 2)
  double* hout = nullptr;       // V^H y (modes 0, 1) or ||y||^2 (mode 2)
  const int* halt = nullptr;    // device flag: non-zero turns the launch into a no-op
  const char* family = nullptr; // profiling family override
  MailPull pull;                // P > 1: hin is the sum of the per-rank partials in the mailbox
  MailPush push;                // P > 1: the result goes to every peer's mailbox instead of hout
  // Pythagorean norm (row-partitioned fast path): UPDATE_DOT with norm_trick also reduces ||y||^2 as entry ncols of
  // hout; UPDATE_NORM with norm_trick then needs no reduction of its own: hout[0] = hin[ncols] - sum |hin_j|^2
  // (= ||y - V hin||^2 for orthonormal V), computed by one CTA from the already reduced coefficients.
  int norm_trick = 0;
  // Guard of the Pythagorean norm: when beta^2 <= norm_guard * ||y||^2 (cancellation: the relative error of beta^2 is
  // eps ||y||^2 / beta^2) the UPDATE_NORM launch raises the sticky halt flag and records retry[0] = 1, retry[1] =
  // retry_tag; the host then reduces the norm explicitly and resumes the chain (krylov.cu).  The decision is taken from
  // values that are bit-identical on every rank.
  int* retry = nullptr;
  int retry_tag = 0;
  double norm_guard = 1e-8;
  SlabPush slab;  // UPDATE_NORM only: n > 0 pushes the output to peer ranks (see SlabPush)
  // The stream operation before this launch does not write the chunk's columns (it is another pass over the same
  // basis): the kernel may fetch its first V tiles while that operation is still running (programmatic dependent launch).
  bool v_stable = false;
};
enum { CGS_DOT = 0, CGS_UPDATE_DOT = 1, CGS_UPDATE_NORM = 2 };
// stand-alone version of the slab push (the first apply of a run has no producing pass): w -> the peers' buffers
int slab_push(cmb_ctx* ctx, const double* w, const SlabPush& slab, const int* halt);
inline int cgs_max_cols(bool cplx) { return cplx ? 64 : 128; }
int cgs_pass(cmb_ctx* ctx, bool cplx, int mode, const CgsPass& a);

// ---- vecops.cu ---------------------------------------------------------------------------------------
// out[0] = sum conj(a_i) b_i (complex: out[0]=re, out[1]=im) over ld doubles; deterministic two-stage
int vec_dot(cmb_ctx* ctx, bool cplx, const double* a, const double* b, int64_t ld, double* out, const int* halt);
// y = x * (1/sqrt(nrm2[0]))
int vec_scale_rsqrt(cmb_ctx* ctx, const double* x, const double* nrm2, double* y, int64_t ld, const int* halt);
// y += shift * x   (complex shift: sr + i si)
int vec_axpy_shift(cmb_ctx* ctx, bool cplx, double sr, double si, const double* x, double* y, int64_t ld,
                   const int* halt);
// x *= (fr + i fi) / sqrt(nrm2[0])  (final Ritz-vector normalisation + phase)
int vec_scale_phase(cmb_ctx* ctx, bool cplx, double* x, const double* nrm2, const double* phase_src, int64_t ld);
// widen a real vector to complex (re, 0)
int vec_real_to_complex(cmb_ctx* ctx, const double* x, double* z, int64_t n);
// finds the index of the first element (scalar index) with |x_i| > 0; writes it to out (INT64_MAX if none)
int vec_first_nonzero(cmb_ctx* ctx, bool cplx, const double* x, int64_t n_scalars, int64_t offset, unsigned long long* out);
int vec_pick_element(cmb_ctx* ctx, const double* x, const unsigned long long* idx, int64_t offset, int64_t n, int es,
                     double* out);

// Per-step scalars every operator-apply kernel reads (device memory, filled by earlier launches):
// nrm2 -> beta = sqrt(nrm2); if beta <= threshold the chain halts (lanczos.hpp:433-437).
struct StepScalars {
  const double* nrm2;  // ||w||^2 (already reduced over ranks) when nrm2_pull.P == 1
  MailPull nrm2_pull;  // P > 1: the per-rank partial norms sit in the mailbox (summed in the prologue)
  double threshold;    // halt when sqrt(nrm2) <= threshold (negative: never)
  int* halt;           // sticky halt flag
  double* beta_slot;   // receives sqrt(nrm2)
  double* alpha_slot;  // receives Re<u_new, v> (complex: 2 doubles), local part
};

}  // namespace cmb
