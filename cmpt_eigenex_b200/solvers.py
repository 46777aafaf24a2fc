"""Python mirror of the reference's solver interface over the C binding (cmpt_b200_solver.h).

Context / DeviceOperator / LanczosEigenSolver / ArnoldiEigenSolver carry the reference's method
names (setMaxIterations, setTolerance, compute, eigenvalues, ... — lanczos.hpp:517-647,
arnoldi.hpp:537-671) so the parity tests read like the reference's samples.  All arithmetic happens
in libcmpt_b200.so on the GPU; this file only marshals arguments.
"""
import ctypes as C
import weakref

import numpy as np

from . import capi
from .capi import check, lib, ptr


class VirtualGroup:
    """nranks virtual ranks on one device (cmpt_b200_debug.h): test vehicle for the multi-rank code paths on a box
    with a single GPU.  Every rank must be driven by its own host thread; see run_virtual_ranks()."""

    def __init__(self, device=0, nranks=2):
        h = C.c_void_p()
        check(lib().cmb_vgroup_create(int(device), int(nranks), C.byref(h)))
        self.h, self.device, self.nranks = h, device, nranks

    def info(self):
        n, g, sm = C.c_int(), C.c_int(), C.c_int()
        check(lib().cmb_vgroup_info(self.h, C.byref(n), C.byref(g), C.byref(sm)))
        return {"nranks": n.value, "green_contexts": bool(g.value), "sms_per_rank": sm.value}

    def close(self):
        if self.h:
            lib().cmb_vgroup_destroy(self.h)
            self.h = None


def run_virtual_ranks(nranks, fn, device=0, timeout=600.0):
    """Runs fn(ctx, comm) on nranks virtual ranks, one host thread each, and returns the list of results.  comm offers
    rank, world, barrier(), gather(obj) -> list over ranks, row_range(n).  The first exception of any rank is re-raised."""
    import threading

    group = VirtualGroup(device, nranks)
    barrier = threading.Barrier(nranks)
    box = [None] * nranks
    results, errors = [None] * nranks, [None] * nranks

    class Comm:
        def __init__(self, rank):
            self.rank, self.world = rank, nranks

        def barrier(self):
            barrier.wait(timeout)

        def gather(self, obj):
            box[self.rank] = obj
            barrier.wait(timeout)
            out = list(box)
            barrier.wait(timeout)
            return out

        def row_range(self, n):
            return (self.rank * n) // nranks, ((self.rank + 1) * n) // nranks

    def work(rank):
        ctx = None
        try:
            ctx = Context(device, virtual=(group, rank))
            results[rank] = fn(ctx, Comm(rank))
        except BaseException as e:  # noqa: BLE001 - re-raised in the caller's thread
            errors[rank] = e
            barrier.abort()
        finally:
            try:
                if ctx is not None:
                    ctx.close()
            except Exception:
                pass

    threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    group_info = group.info()
    alive = [t.is_alive() for t in threads]
    if not any(alive):
        group.close()
    first = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)] or [e for e in errors if e is not None]
    if first:
        others = ["rank %d: %s: %s" % (r, type(e).__name__, e) for r, e in enumerate(errors)
                  if e is not None and e is not first[0] and not isinstance(e, threading.BrokenBarrierError)]
        if others:
            raise RuntimeError("%s: %s  [other ranks: %s]" % (type(first[0]).__name__, first[0], " | ".join(others))) from first[0]
        raise first[0]
    if any(alive):
        raise TimeoutError("virtual ranks did not finish within %.0f s" % timeout)
    return results, group_info


class Context:
    """One GPU / one rank (cmb_ctx)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None, virtual=None):
        h = C.c_void_p()
        if virtual is not None:
            group, rank = virtual
            nranks = group.nranks
            check(lib().cmb_ctx_create_virtual(group.h, int(rank), C.byref(h)))
        elif nranks == 1:
            check(lib().cmb_ctx_create(int(device), C.byref(h)))
        else:
            idbuf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            check(lib().cmb_ctx_create_dist(int(device), int(rank), int(nranks), idbuf, C.byref(h)))
        self.h = h
        self.rank, self.nranks, self.device = rank, nranks, device
        # operators and solvers that use this context: they must be destroyed before it
        self._ops = weakref.WeakSet()
        self._solvers = weakref.WeakSet()

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        check(lib().cmb_nccl_unique_id(buf))
        return bytes(buf)

    def sync(self):
        check(lib().cmb_ctx_sync(self.h))

    def timer_start(self):
        check(lib().cmb_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        check(lib().cmb_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib().cmb_ctx_launch_count(self.h))

    def profile(self, enable=True):
        check(lib().cmb_ctx_profile(self.h, 1 if enable else 0))

    def profile_get(self, family):
        ms, n = C.c_double(), C.c_uint64()
        check(lib().cmb_ctx_profile_get(self.h, family.encode(), C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def flush_l2(self):
        check(lib().cmb_ctx_flush_l2(self.h))

    def set_spin_timeout(self, seconds):
        check(lib().cmb_ctx_set_spin_timeout(self.h, float(seconds)))

    def close(self):
        if self.h:
            # a context that a peer timeout marked dead makes the solvers' and operators' destructors report errors:
            # the context itself must still be destroyed, and only once
            try:
                for s in list(self._solvers):
                    try:
                        s.close()
                    except Exception:
                        pass
                for o in list(self._ops):
                    try:
                        o.close()
                    except Exception:
                        pass
            finally:
                h, self.h = self.h, None
                lib().cmb_ctx_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceOperator:
    """Operator resident in HBM (cmb_op): the device replacement of the matmul callback."""

    def __init__(self, ctx, handle, keep=None):
        self.ctx, self.h, self._keep = ctx, handle, keep
        ctx._ops.add(self)

    @staticmethod
    def from_csr(ctx, rowptr, col, val, n_global=None, row_begin=0):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        code = capi.dtype_code(val.dtype)
        val = np.ascontiguousarray(val, dtype=capi.np_dtype(code))
        nloc = rowptr.size - 1
        n_global = nloc if n_global is None else n_global
        if nloc < 0 or rowptr[0] != 0 or col.size < rowptr[-1] or val.size < rowptr[-1]:
            raise ValueError("inconsistent CSR arrays: rowptr must start at 0 and col/val must hold rowptr[-1] entries")
        h = C.c_void_p()
        check(lib().cmb_op_csr_create(ctx.h, code, n_global, row_begin, row_begin + nloc, ptr(rowptr), ptr(col),
                                      ptr(val), C.byref(h)))
        return DeviceOperator(ctx, h)

    @staticmethod
    def from_dense(ctx, a):
        code = capi.dtype_code(a.dtype)
        a = np.ascontiguousarray(a, dtype=capi.np_dtype(code))
        n = a.shape[0]
        h = C.c_void_p()
        check(lib().cmb_op_dense_create(ctx.h, code, n, 0, n, ptr(a), C.byref(h)))
        return DeviceOperator(ctx, h)

    @staticmethod
    def heisenberg(ctx, L, J=1.0, pbc=True, dtype=np.float64):
        h = C.c_void_p()
        check(lib().cmb_op_heisenberg_create(ctx.h, capi.dtype_code(dtype), int(L), float(J), int(bool(pbc)),
                                             C.byref(h)))
        return DeviceOperator(ctx, h)

    @staticmethod
    def linear(ctx, ops, coefs):
        """sum_i coefs[i] * ops[i] as one operator that stays in HBM (cmb_op_linear_create); the terms must outlive it."""
        dt = ops[0].dtype
        c = np.ascontiguousarray(coefs, dtype=dt)
        arr = (C.c_void_p * len(ops))(*[o.h for o in ops])
        h = C.c_void_p()
        check(lib().cmb_op_linear_create(ctx.h, len(ops), arr, ptr(c), C.byref(h)))
        return DeviceOperator(ctx, h, keep=list(ops))

    @staticmethod
    def product(ctx, outer, inner):
        """outer * inner (inner acts first) as one operator that stays in HBM (cmb_op_product_create)."""
        h = C.c_void_p()
        check(lib().cmb_op_product_create(ctx.h, outer.h, inner.h, C.byref(h)))
        return DeviceOperator(ctx, h, keep=[outer, inner])

    @staticmethod
    def from_callback(ctx, fn, n, dtype=np.float64):
        cb, code = _make_callback(fn, n, dtype)
        h = C.c_void_p()
        check(lib().cmb_op_callback_create(ctx.h, code, n, cb, None, C.byref(h)))
        return DeviceOperator(ctx, h, keep=cb)

    @property
    def height(self):
        return int(lib().cmb_op_height(self.h))

    @property
    def rows(self):
        return int(lib().cmb_op_rows(self.h))

    @property
    def dtype(self):
        return capi.np_dtype(lib().cmb_op_dtype(self.h))

    @property
    def bytes(self):
        return float(lib().cmb_op_bytes(self.h))

    def sell_stats(self):
        """CSR operators: (nnz, stored entries incl. padding, rows sorted inside 1024-row windows?)"""
        nnz, padded, srt = C.c_longlong(), C.c_longlong(), C.c_int()
        check(lib().cmb_debug_op_sell_stats(self.h, C.byref(nnz), C.byref(padded), C.byref(srt)))
        return int(nnz.value), int(padded.value), bool(srt.value)

    def apply(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.empty_like(x)
        check(lib().cmb_op_apply_host(self.h, ptr(x), ptr(y)))
        return y

    def close(self):
        if self.h:
            lib().cmb_op_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _make_callback(fn, n, dtype):
    """Wrap y = fn(x) (numpy) into the reference's matmul signature (const Scalar*, Scalar*)."""
    code = capi.dtype_code(dtype)
    dt = capi.np_dtype(code)
    width = 2 if code == capi.CMB_C64 else 1

    def tramp(pin, pout, _user):
        x = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_double)), shape=(n * width,)).view(dt)
        y = np.ctypeslib.as_array(C.cast(pout, C.POINTER(C.c_double)), shape=(n * width,)).view(dt)
        y[:] = fn(x)

    return capi.MATMUL_FN(tramp), code


class _Solver:
    kind = None
    unlimited = -1

    def __init__(self, dtype=np.float64):
        self.code = capi.dtype_code(dtype)
        self.dtype = capi.np_dtype(self.code)
        h = C.c_void_p()
        check(lib().cmbs_create(self.kind, self.code, C.byref(h)))
        self.h = h
        self._op = None
        self._cb = None

    # ---- settings (names follow the reference) ----
    def setMatrixMultiplication(self, op, height=None):
        """op: DeviceOperator (device path) or a Python callable y = f(x) with `height` (legacy path)."""
        if isinstance(op, DeviceOperator):
            check(lib().cmbs_set_operator(self.h, op.ctx.h, op.h))
            self._op = op
            op.ctx._solvers.add(self)
        else:
            cb, _ = _make_callback(op, int(height), self.dtype)
            check(lib().cmbs_set_callback(self.h, int(height), cb, None))
            self._cb = cb
        return self

    def _seti(self, k, v):
        check(lib().cmbs_set_int(self.h, k.encode(), int(v)))
        return self

    def _setr(self, k, v):
        check(lib().cmbs_set_real(self.h, k.encode(), float(v)))
        return self

    def _geti(self, k):
        v = C.c_int64()
        check(lib().cmbs_get_int(self.h, k.encode(), C.byref(v)))
        return int(v.value)

    def _getr(self, k):
        v = C.c_double()
        check(lib().cmbs_get_real(self.h, k.encode(), C.byref(v)))
        return v.value

    def setMinIterations(self, v):
        return self._seti("minIterations", v)

    def setMaxIterations(self, v):
        return self._seti("maxIterations", v)

    def setMaxEigenvalues(self, v):
        return self._seti("maxEigenvalues", v)

    def setComputeEigenvectorsOn(self, v):
        return self._seti("computeEigenvectorsOn", 1 if v else 0)

    def setReserveSize(self, v):
        return self._seti("reserveSize", v)

    def setTolerance(self, v):
        return self._setr("tolerance", v)

    def setThreshold(self, v):
        return self._setr("threshold", v)

    def setEigenvalueShift(self, v):
        if isinstance(v, complex) and self.kind in (capi.CMBS_ARNOLDI, capi.CMBS_THICK_RESTART_ARNOLDI):
            check(lib().cmbs_set_complex(self.h, b"eigenvalueShift", v.real, v.imag))
            return self
        return self._setr("eigenvalueShift", float(np.real(v)))

    def setIndicesForConvergence(self, idx):
        a = np.ascontiguousarray(idx, dtype=np.int64)
        check(lib().cmbs_set_indices_for_convergence(self.h, ptr(a), a.size))
        return self

    def setInitialVector(self, v=None):
        if v is None:
            check(lib().cmbs_set_initial_vector(self.h, None, 0))
        else:
            a = np.ascontiguousarray(v, dtype=self.dtype)
            check(lib().cmbs_set_initial_vector(self.h, ptr(a), a.size))
        return self

    def setOrthogonalizingVectors(self, vecs):
        if len(vecs) == 0:
            check(lib().cmbs_set_orthogonalizing_vectors(self.h, 0, None, 0))
            return self
        m = np.asfortranarray(np.stack([np.asarray(v, dtype=self.dtype) for v in vecs], axis=1))
        check(lib().cmbs_set_orthogonalizing_vectors(self.h, m.shape[1], ptr(m), m.shape[0]))
        return self

    # ---- run ----
    def compute(self):
        check(lib().cmbs_compute(self.h))
        return 0

    def continueToCompute(self):
        check(lib().cmbs_continue_to_compute(self.h))
        return 0

    def clear(self):
        check(lib().cmbs_clear(self.h))

    def clearComputedData(self):
        check(lib().cmbs_clear_computed_data(self.h))

    # ---- results ----
    def iterations(self):
        return self._geti("iterations")

    def matrixHeight(self):
        return self._geti("matrixHeight")

    def nvectors(self):
        return self._geti("nvectors")

    def eigenvectors(self, copy=True):
        p, r, c = C.c_void_p(), C.c_int64(), C.c_int64()
        check(lib().cmbs_eigenvectors_ptr(self.h, C.byref(p), C.byref(r), C.byref(c)))
        edt = self._vec_dtype()
        if r.value * c.value == 0:
            return np.zeros((r.value, c.value), dtype=edt, order="F")
        nbytes = r.value * c.value * np.dtype(edt).itemsize
        buf = (C.c_char * nbytes).from_address(p.value)
        a = np.frombuffer(buf, dtype=edt).reshape((r.value, c.value), order="F")
        return a.copy(order="F") if copy else a

    def ritzResiduals(self):
        out = np.empty(self._geti("neigenvalues"))
        if out.size:
            check(lib().cmbs_get_ritz_residuals(self.h, ptr(out)))
        return out

    def basisVector(self, k):
        out = np.empty(self._geti("localHeight"), dtype=self.dtype)
        check(lib().cmbs_get_basis_vector(self.h, k, ptr(out)))
        return out

    def log(self):
        out = []
        buf = C.create_string_buffer(512)
        for i in range(self._geti("nlog")):
            check(lib().cmbs_get_log_line(self.h, i, buf, 512))
            out.append(buf.value.decode())
        return out

    def hasWARN(self):
        return self._geti("hasWARN")

    def hasERROR(self):
        return self._geti("hasERROR")

    def deviceBytes(self):
        return float(lib().cmbs_device_bytes(self.h))

    def close(self):
        if self.h:
            lib().cmbs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LanczosEigenSolver(_Solver):
    """cmpt::EigenEx::LanczosEigenSolver<Scalar> (lanczos.hpp:468-927) on the GPU."""

    kind = capi.CMBS_LANCZOS

    def _vec_dtype(self):
        return self.dtype

    def setReorthogonalizeInterval(self, v):
        return self._seti("reorthogonalizeInterval", v)

    def eigenvalues(self):
        out = np.empty(self._geti("neigenvalues"))
        if out.size:
            check(lib().cmbs_get_eigenvalues(self.h, ptr(out)))
        return out

    def alpha(self):
        return self._ab()[0]

    def beta(self):
        return self._ab()[1]

    def _ab(self):
        a, b = np.empty(self._geti("nalpha")), np.empty(self._geti("nbeta"))
        check(lib().cmbs_get_alpha_beta(self.h, ptr(a) if a.size else ptr(np.empty(1)), ptr(b) if b.size else ptr(np.empty(1))))
        return a, b

    def es_tri_eigenvectors(self):
        r, c = C.c_int64(), C.c_int64()
        check(lib().cmbs_get_small_eigenvectors(self.h, None, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), dtype=self.dtype, order="F")
        if out.size:
            check(lib().cmbs_get_small_eigenvectors(self.h, ptr(out), C.byref(r), C.byref(c)))
        return out

    def convergenceLog(self, index):
        n = C.c_int64()
        check(lib().cmbs_get_convergence_log(self.h, index, None, C.byref(n)))
        out = np.empty(n.value)
        if n.value:
            check(lib().cmbs_get_convergence_log(self.h, index, ptr(out), C.byref(n)))
        return out

    # LanczosExponentialSolver (lanczos.hpp:1002-1164)
    def expSolveWithLanczos(self, x):
        """compute(), then exp(x A) applied to the solver's initial vector (local slab)."""
        out = np.empty(self._geti("localHeight"), dtype=self.dtype)
        x = complex(x)
        check(lib().cmbs_exp_solve_with_lanczos(self.h, x.real, x.imag, ptr(out)))
        return out

    def expSolveWithTaylor(self, x, radius, vin, auto_division=True):
        vin = np.ascontiguousarray(vin, dtype=self.dtype)
        out = np.empty_like(vin)
        x = complex(x)
        check(lib().cmbs_exp_solve_with_taylor(self.h, x.real, x.imag, float(radius), int(bool(auto_division)),
                                               ptr(vin), ptr(out)))
        return out


class ThickRestartLanczos(_Solver):
    """cmpt::EigenEx::ThickRestartLanczos<Scalar> (include/cmpt/eigen_ex/thick_restart.hpp; additive): the lowest
    `wanted` eigenpairs with at most `maxBasis` Lanczos vectors on the device."""

    kind = capi.CMBS_THICK_RESTART

    def _vec_dtype(self):
        return self.dtype

    def setWanted(self, v):
        return self._seti("wanted", v)

    def setMaxBasis(self, v):
        return self._seti("maxBasis", v)

    def setKeep(self, v):
        return self._seti("keep", v)

    def setMaxRestarts(self, v):
        return self._seti("maxRestarts", v)

    def eigenvalues(self):
        out = np.empty(self._geti("neigenvalues"))
        if out.size:
            check(lib().cmbs_get_eigenvalues(self.h, ptr(out)))
        return out

    def residuals(self):
        return self.ritzResiduals()

    def restarts(self):
        return self._geti("restarts")

    def operatorApplications(self):
        return self._geti("operatorApplications")

    def converged(self):
        return self._geti("converged")


class ThickRestartArnoldi(ThickRestartLanczos):
    """cmpt::EigenEx::ThickRestartArnoldi<Scalar> (include/cmpt/eigen_ex/arnoldi_restart.hpp; additive, Krylov-Schur):
    the `wanted` eigenpairs of a general operator with at most `maxBasis` Arnoldi vectors on the device.  Eigenvalues and
    eigenvectors are complex for both Scalars."""

    kind = capi.CMBS_THICK_RESTART_ARNOLDI
    LARGEST_MAGNITUDE, LARGEST_REAL, SMALLEST_REAL = 0, 1, 2

    def _vec_dtype(self):
        return np.complex128

    def setWhich(self, v):
        return self._seti("which", v)

    def eigenvalues(self):
        out = np.empty(self._geti("neigenvalues"), dtype=np.complex128)
        if out.size:
            check(lib().cmbs_get_eigenvalues(self.h, ptr(out)))
        return out


class ArnoldiEigenSolver(_Solver):
    """cmpt::EigenEx::ArnoldiEigenSolver<Scalar> (arnoldi.hpp:444-1027) on the GPU."""

    kind = capi.CMBS_ARNOLDI

    def _vec_dtype(self):
        return np.complex128

    def eigenvalues(self):
        out = np.empty(self._geti("neigenvalues"), dtype=np.complex128)
        if out.size:
            check(lib().cmbs_get_eigenvalues(self.h, ptr(out)))
        return out

    def computeWithRestarts(self, cycles):
        check(lib().cmbs_compute_with_restarts(self.h, int(cycles)))
        return 0

    def hessenbergMatrix(self):
        m = self._geti("hessenbergSize")
        out = np.zeros((m, m), dtype=self.dtype, order="F")
        if m:
            check(lib().cmbs_get_hessenberg(self.h, ptr(out)))
        return out

    def residue(self):
        v = C.c_double()
        check(lib().cmbs_get_residue(self.h, C.byref(v)))
        return v.value

    def eigenvectors_h(self):
        r, c = C.c_int64(), C.c_int64()
        check(lib().cmbs_get_small_eigenvectors(self.h, None, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), dtype=np.complex128, order="F")
        if out.size:
            check(lib().cmbs_get_small_eigenvectors(self.h, ptr(out), C.byref(r), C.byref(c)))
        return out

    def convergenceLog(self, index):
        n = C.c_int64()
        check(lib().cmbs_get_convergence_log(self.h, index, None, C.byref(n)))
        out = np.empty(n.value, dtype=np.complex128)
        if n.value:
            check(lib().cmbs_get_convergence_log(self.h, index, ptr(out), C.byref(n)))
        return out


def host_tridiagonal_eigen(alpha, beta, vectors=True):
    """The product's host tridiagonal solver (detail/tridiag_eigen.hpp); no GPU needed."""
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    b = np.ascontiguousarray(beta, dtype=np.float64)
    n = a.size
    w = np.empty(n)
    z = np.empty((n, n), order="F") if vectors else None
    check(lib().cmbs_host_tridiagonal_eigen(n, ptr(a), ptr(b) if b.size else None, ptr(w), ptr(z) if vectors else None))
    return (w, z) if vectors else w


def host_symmetric_eigen(a):
    """The product's host Jacobi solver for small dense symmetric matrices (detail/symmetric_eigen.hpp): the projected
    matrix of the thick-restart driver; no GPU needed."""
    af = np.asfortranarray(a, dtype=np.float64)
    n = af.shape[0]
    w = np.empty(n)
    z = np.empty((n, n), order="F")
    check(lib().cmbs_host_symmetric_eigen(n, ptr(af), ptr(w), ptr(z)))
    return w, z


def host_general_eigen(a, vectors=True):
    """The product's host solver for a general complex matrix (Householder reduction + Hessenberg QR); no GPU needed."""
    ac = np.asfortranarray(a, dtype=np.complex128)
    n = ac.shape[0]
    w = np.empty(n, dtype=np.complex128)
    v = np.empty((n, n), dtype=np.complex128, order="F") if vectors else None
    check(lib().cmbs_host_general_eigen(n, ptr(ac), ptr(w), ptr(v) if vectors else None))
    return (w, v) if vectors else w


def host_hessenberg_eigen(h, vectors=True):
    """The product's host Hessenberg solver (detail/hessenberg_eigen.hpp); no GPU needed."""
    hc = np.asfortranarray(h, dtype=np.complex128)
    n = hc.shape[0]
    w = np.empty(n, dtype=np.complex128)
    v = np.empty((n, n), dtype=np.complex128, order="F") if vectors else None
    check(lib().cmbs_host_hessenberg_eigen(n, ptr(hc), ptr(w), ptr(v) if vectors else None))
    return (w, v) if vectors else w
