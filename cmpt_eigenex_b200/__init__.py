"""cmpt-eigenex on B200: Lanczos / Arnoldi eigensolvers driving hand-written sm_100a kernels.

Layout
  csrc/          CUDA kernels + the C-ABI (built in-tree to lib/libcmpt_b200.so; headers in include/)
  capi.py        ctypes loader of libcmpt_b200.so (fails loudly when it is missing; no CPU fallback)
  solvers.py     Python mirror of the reference's solver classes over the C binding
  synthetic.py   BASELINE operators / start vectors (numpy, no CUDA)
  dist.py        one-process-per-GPU bootstrap (torch.distributed carries the NCCL id)

The C++ drop-in headers live in include/cmpt/eigen_ex/.
"""
from . import capi, synthetic  # noqa: F401
from .solvers import (ArnoldiEigenSolver, Context, DeviceOperator, LanczosEigenSolver,  # noqa: F401
                      ThickRestartArnoldi, ThickRestartLanczos, VirtualGroup, run_virtual_ranks, host_general_eigen, host_hessenberg_eigen, host_symmetric_eigen, host_tridiagonal_eigen)

__all__ = ["capi", "synthetic", "Context", "DeviceOperator", "LanczosEigenSolver", "ArnoldiEigenSolver", "ThickRestartLanczos", "ThickRestartArnoldi", "VirtualGroup", "run_virtual_ranks",
           "host_tridiagonal_eigen", "host_hessenberg_eigen", "host_general_eigen", "host_symmetric_eigen"]
