// op.cuh — internal operator interface (the device replacement of matrixMultiplication_).
#pragma once
#include "kernels.cuh"

struct cmb_op {
  cmb_ctx* ctx = nullptr;
  int dtype = CMB_F64;
  bool cplx = false;
  int64_t n_global = 0, row_begin = 0, n_local = 0;
  double bytes = 0.0;  // algorithmic bytes of one local apply
  const char* family = "op";
  virtual ~cmb_op() {}
  // Fused step: u = w / sqrt(nrm2) -> ucol ; v = (A + shift) u ; alpha_slot = sum conj(u_i) v_i (local part).
  // w, ucol, v are padded device vectors (doubles).  All launches go to ctx->stream.
  virtual int apply(const double* w, double* ucol, double* v, double shift_re, double shift_im,
                    const cmb::StepScalars& sc) = 0;
  // Fused compute + exchange (row-partitioned matrix-free operators): when the kernel that PRODUCES w can also push the
  // parts of w the partner ranks need, the operator describes the destinations here and the next apply(w, ...) skips
  // its own exchange.  Returns false when the operator has nothing to push (the default).
  virtual bool slab_push_begin(const double* w, cmb::SlabPush* out) {
    (void)w;
    (void)out;
    return false;
  }
  // forget a push announced by slab_push_begin (the chain was halted, or w has been rewritten since)
  virtual void slab_push_cancel() {}
};
