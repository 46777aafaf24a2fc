"""Per-column-count bandwidth of the three Gram-Schmidt passes (diagnostic; run on the GPU box)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import capi
from cmpt_eigenex_b200.capi import check, lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
cplx = len(sys.argv) > 2 and sys.argv[2] == "z"
cols = [1, 2, 3, 4, 6, 8, 9, 12, 16, 17, 24, 32, 33, 40, 48, 56, 64] + ([] if cplx else [65, 72, 74, 80, 88, 96, 100, 104, 112, 128])
ctx = pkg.Context(0)
K = C.c_void_p()
check(lib().cmb_krylov_create(ctx.h, 1 if cplx else 0, n, 0, n, 128, C.byref(K)))
s = 16 if cplx else 8
print("n=%d %s   c : GB/s dot / update_dot / update_norm   (ms)" % (n, "complex" if cplx else "real"))
for c in cols:
    out = []
    for mode in (0, 1, 2):
        ms = C.c_double()
        check(lib().cmb_debug_cgs_pass(K, mode, c, 5, C.byref(ms)))
        bytes_ = (c + (1 if mode == 0 else 2)) * n * s
        out.append((bytes_ / ms.value / 1e6, ms.value))
    print("%4d : %7.0f %7.0f %7.0f    (%.3f %.3f %.3f)" % (c, out[0][0], out[1][0], out[2][0], out[0][1], out[1][1], out[2][1]))
lib().cmb_krylov_destroy(K)
