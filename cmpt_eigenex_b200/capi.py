"""ctypes loader of libcmpt_b200.so (the C-ABI of include/cmpt_b200.h and cmpt_b200_solver.h).

The library is the product: there is no Python or CPU fallback.  Loading fails loudly when the
shared object has not been built (run `python -c "import __graft_entry__ as g; g.build()"`), and
every compute entry point fails with CMB_ERR_NO_DEVICE when no B200 is visible.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcmpt_b200.so")

CMB_F64, CMB_C64 = 0, 1
CMBS_LANCZOS, CMBS_ARNOLDI, CMBS_THICK_RESTART, CMBS_THICK_RESTART_ARNOLDI = 0, 1, 2, 3
CMB_OK = 0
CMB_ERR_NO_DEVICE = -6
STEP_OK, STEP_BREAKDOWN, STEP_NOSTART, STEP_FULL = 0, 1, 2, 4

MATMUL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)

_lib = None


class CmbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libcmpt_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """The loaded library; raises if it is missing (never falls back to anything else)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libcmpt_b200.so is not built (%s missing): run __graft_entry__.build(); "
                "there is no CPU fallback" % LIB_PATH
            )
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i64, i32, dbl, cp = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_char_p
    P = C.POINTER
    sig = {
        "cmb_version": (cp, []),
        "cmb_last_error": (cp, []),
        "cmb_device_count": (i32, [P(i32)]),
        "cmb_ctx_create": (i32, [i32, P(vp)]),
        "cmb_nccl_unique_id": (i32, [vp]),
        "cmb_ctx_create_dist": (i32, [i32, i32, i32, vp, P(vp)]),
        "cmb_ctx_destroy": (i32, [vp]),
        "cmb_ctx_rank": (i32, [vp]),
        "cmb_ctx_nranks": (i32, [vp]),
        "cmb_ctx_sync": (i32, [vp]),
        "cmb_ctx_timer_start": (i32, [vp]),
        "cmb_ctx_timer_stop": (i32, [vp, P(dbl)]),
        "cmb_ctx_launch_count": (C.c_uint64, [vp]),
        "cmb_ctx_profile": (i32, [vp, i32]),
        "cmb_ctx_profile_get": (i32, [vp, cp, P(dbl), P(C.c_uint64)]),
        "cmb_ctx_flush_l2": (i32, [vp]),
        "cmb_op_csr_create": (i32, [vp, i32, i64, i64, i64, vp, vp, vp, P(vp)]),
        "cmb_partition_begin": (i64, [i64, i32, i32]),
        "cmb_plan_halo": (i32, [i64, i32, i32, i64, vp, vp, P(i64), vp, vp, i64]),
        "cmb_op_dense_create": (i32, [vp, i32, i64, i64, i64, vp, P(vp)]),
        "cmb_op_heisenberg_create": (i32, [vp, i32, i32, dbl, i32, P(vp)]),
        "cmb_debug_heisenberg_virtual": (i32, [vp, i32, i32, dbl, i32, i32, vp, vp]),
        "cmb_heisenberg_plan": (i32, [i32, i32, i32, i32, P(i32), P(i32), P(i32), P(i32), P(i32)]),
        "cmb_op_callback_create": (i32, [vp, i32, i64, MATMUL_FN, vp, P(vp)]),
        "cmb_op_linear_create": (i32, [vp, i64, vp, vp, P(vp)]),
        "cmb_op_product_create": (i32, [vp, vp, vp, P(vp)]),
        "cmb_op_destroy": (i32, [vp]),
        "cmb_op_context": (vp, [vp]),
        "cmb_op_row_begin": (i64, [vp]),
        "cmb_op_rows": (i64, [vp]),
        "cmb_op_height": (i64, [vp]),
        "cmb_op_dtype": (i32, [vp]),
        "cmb_op_bytes": (dbl, [vp]),
        "cmb_op_apply_host": (i32, [vp, vp, vp]),
        "cmb_krylov_create": (i32, [vp, i32, i64, i64, i64, i64, P(vp)]),
        "cmb_krylov_destroy": (i32, [vp]),
        "cmb_krylov_clear": (i32, [vp]),
        "cmb_krylov_set_deflation": (i32, [vp, i64, vp, i64]),
        "cmb_krylov_start": (i32, [vp, vp, dbl, P(i32)]),
        "cmb_krylov_restart": (i32, [vp, dbl, P(i32)]),
        "cmb_krylov_ncols": (i64, [vp]),
        "cmb_krylov_rows": (i64, [vp]),
        "cmb_krylov_get_col": (i32, [vp, i64, vp]),
        "cmb_lanczos_step": (i32, [vp, vp, dbl, i64, dbl, P(dbl), P(dbl), P(i32)]),
        "cmb_lanczos_run": (i32, [vp, vp, dbl, i64, dbl, i64, vp, vp, P(i64), P(i32)]),
        "cmb_lanczos_residual_norm": (i32, [vp, P(dbl)]),
        "cmb_lanczos_thick_restart": (i32, [vp, P(dbl), i64, i64, i64]),
        "cmb_arnoldi_thick_restart": (i32, [vp, vp, i64, i64, i64]),
        "cmb_arnoldi_run": (i32, [vp, vp, vp, dbl, i64, vp, i64, vp, P(i64), P(i32)]),
        "cmb_arnoldi_step": (i32, [vp, vp, vp, dbl, vp, P(dbl), P(i32)]),
        "cmb_krylov_ritz_vectors": (i32, [vp, i32, vp, i64, i64, i64, vp, i64]),
        "cmb_krylov_bytes": (dbl, [vp]),
        "cmb_krylov_project": (i32, [vp, vp, vp]),
        "cmb_krylov_combine": (i32, [vp, vp, i64, vp]),
        "cmb_debug_cgs_pass": (i32, [vp, i32, i32, i32, P(dbl)]),
        "cmb_debug_op_exchange_count": (i32, [vp, P(C.c_longlong)]),
        "cmb_debug_op_sell_stats": (i32, [vp, P(C.c_longlong), P(C.c_longlong), P(i32)]),
        "cmb_vgroup_create": (i32, [i32, i32, P(vp)]),
        "cmb_vgroup_destroy": (i32, [vp]),
        "cmb_vgroup_info": (i32, [vp, P(i32), P(i32), P(i32)]),
        "cmb_ctx_create_virtual": (i32, [vp, i32, P(vp)]),
        "cmb_ctx_set_spin_timeout": (i32, [vp, dbl]),
        # solver binding
        "cmbs_create": (i32, [i32, i32, P(vp)]),
        "cmbs_destroy": (i32, [vp]),
        "cmbs_set_operator": (i32, [vp, vp, vp]),
        "cmbs_set_callback": (i32, [vp, i64, MATMUL_FN, vp]),
        "cmbs_set_int": (i32, [vp, cp, i64]),
        "cmbs_set_real": (i32, [vp, cp, dbl]),
        "cmbs_set_complex": (i32, [vp, cp, dbl, dbl]),
        "cmbs_get_int": (i32, [vp, cp, P(i64)]),
        "cmbs_get_real": (i32, [vp, cp, P(dbl)]),
        "cmbs_set_indices_for_convergence": (i32, [vp, vp, i64]),
        "cmbs_set_initial_vector": (i32, [vp, vp, i64]),
        "cmbs_set_orthogonalizing_vectors": (i32, [vp, i64, vp, i64]),
        "cmbs_compute": (i32, [vp]),
        "cmbs_continue_to_compute": (i32, [vp]),
        "cmbs_compute_with_restarts": (i32, [vp, i64]),
        "cmbs_clear": (i32, [vp]),
        "cmbs_clear_computed_data": (i32, [vp]),
        "cmbs_get_eigenvalues": (i32, [vp, vp]),
        "cmbs_eigenvectors_ptr": (i32, [vp, P(vp), P(i64), P(i64)]),
        "cmbs_get_ritz_residuals": (i32, [vp, vp]),
        "cmbs_get_alpha_beta": (i32, [vp, vp, vp]),
        "cmbs_get_hessenberg": (i32, [vp, vp]),
        "cmbs_get_residue": (i32, [vp, P(dbl)]),
        "cmbs_get_small_eigenvectors": (i32, [vp, vp, P(i64), P(i64)]),
        "cmbs_get_basis_vector": (i32, [vp, i64, vp]),
        "cmbs_get_log_line": (i32, [vp, i64, C.c_char_p, i64]),
        "cmbs_get_convergence_log": (i32, [vp, i64, vp, P(i64)]),
        "cmbs_device_bytes": (dbl, [vp]),
        "cmbs_exp_solve_with_lanczos": (i32, [vp, dbl, dbl, vp]),
        "cmbs_exp_solve_with_taylor": (i32, [vp, dbl, dbl, dbl, i32, vp, vp]),
        "cmbs_host_tridiagonal_eigen": (i32, [i64, vp, vp, vp, vp]),
        "cmbs_host_hessenberg_eigen": (i32, [i64, vp, vp, vp]),
        "cmbs_host_general_eigen": (i32, [i64, vp, vp, vp]),
        "cmbs_host_symmetric_eigen": (i32, [i64, vp, vp, vp]),
        "cmb_host_alloc": (i32, [C.c_size_t, P(vp)]),
        "cmb_host_free": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    L._declared = sorted(sig)


def declared_symbols():
    lib()
    return list(_lib._declared)


def check(rc):
    if rc != CMB_OK:
        raise CmbError(rc, lib().cmb_last_error().decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dtype_code(dt):
    return CMB_C64 if np.issubdtype(np.dtype(dt), np.complexfloating) else CMB_F64


def np_dtype(code):
    return np.complex128 if code == CMB_C64 else np.float64


def device_count():
    n = C.c_int(0)
    check(lib().cmb_device_count(C.byref(n)))
    return n.value


class PinnedBuffer:
    """numpy view of pinned host memory (cmb_host_alloc) for full-rate PCIe staging."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape))
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib().cmb_host_alloc(max(nbytes, 16), C.byref(p)))
        self._p = p
        buf = (C.c_char * max(nbytes, 16)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            lib().cmb_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
