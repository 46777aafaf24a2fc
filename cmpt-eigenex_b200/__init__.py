"""cmpt-eigenex on B200: Lanczos / Arnoldi eigensolvers driving hand-written sm_100a kernels.

Layout
  csrc/          CUDA kernels + the C-ABI (libcmpt_b200.so; header include/cmpt_b200.h)
  synthetic.py   BASELINE operators / start vectors (numpy, no CUDA)
  capi.py        ctypes loader of libcmpt_b200.so (fails loudly when it is missing)
  solvers.py     Python mirror of the reference's solver classes over the C binding

The C++ drop-in headers live in include/cmpt/eigen_ex/.
"""
from . import synthetic  # noqa: F401

__all__ = ["synthetic"]
