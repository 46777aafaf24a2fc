"""ctypes binding of oracle/_ref/libref.so — the UNMODIFIED reference classes (TEST INFRASTRUCTURE ONLY).

libref.so is /root/reference/include/cmpt/eigen_ex/{lanczos,arnoldi}.hpp compiled as they are against
oracle/eigen_shim/ behind the C-ABI of oracle/ref_harness.cpp (recipe: oracle/Makefile, target `ref`).
The classes below expose the same Python surface as oracle/core.py (step level) and
oracle/reference_solvers.py (solver level), so a test can run the restatement and the reference itself
through identical code.  Operators are oracle/core.py `Operator` handles: the reference's MatMulFunction
calls the very same host routine the restatement uses, so any difference comes from the Krylov code alone.

Importers: tests/, tests/golden/make_ref_golden.py, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import core

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE = os.environ.get("CMPT_REFERENCE_ROOT", "/root/reference")
_LIB = None
_DT = {"d": np.float64, "z": np.complex128}

APPLY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)


def reference_present():
    return os.path.exists(os.path.join(REFERENCE, "include", "cmpt", "eigen_ex", "lanczos.hpp"))


def available():
    """True when libref.so can be loaded: either it is prebuilt (GPU box) or the reference tree is here."""
    return os.path.exists(_SO) or reference_present()


def build(force=False):
    """(Re)build oracle/_ref/ with the committed Makefile when the reference tree is present."""
    if reference_present():
        env = dict(os.environ)
        env.pop("CXX", None)
        cmd = ["make", "-C", _HERE, "ref", "REFERENCE=" + REFERENCE]
        if force:
            cmd.insert(1, "-B")
        subprocess.check_call(cmd, env=env, stdout=subprocess.DEVNULL)
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/libref.so is missing and %s is not present to build it from" % REFERENCE)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _declare(_LIB)
    return _LIB


def _declare(L):
    vp, i64, dbl, u32, cint = C.c_void_p, C.c_int64, C.c_double, C.c_uint32, C.c_int
    L.ref_num_threads.restype = cint
    L.ref_set_num_threads.argtypes = [cint]

    def sig(name, restype, argtypes):
        f = getattr(L, name)
        f.restype = restype
        f.argtypes = argtypes

    for p in "dz":
        g = "ref_%s_" % p
        sig(g + "default_vector", None, [i64, vp])
        sig(g + "seeded_vector", None, [u32, i64, vp])
        sig(g + "arnoldi_default_vector", None, [i64, vp])
        for k in ("lanczos", "arnoldi"):
            sig(g + k + "_create", vp, [])
            sig(g + k + "_destroy", None, [vp])
            sig(g + k + "_set_op", None, [vp, vp, vp, i64])
            sig(g + k + "_set_init", None, [vp, vp, i64])
            sig(g + k + "_add_ortho", None, [vp, vp, i64])
            sig(g + k + "_clear_steps", None, [vp])
            sig(g + k + "_step", cint, [vp])
            sig(g + k + "_utmost", cint, [vp])
            sig(g + k + "_iterations", i64, [vp])
            sig(g + k + "_nvectors", i64, [vp])
            sig(g + k + "_get_vector", None, [vp, i64, vp])
        sig(g + "lanczos_set_params", None, [vp, dbl, i64, dbl])
        sig(g + "lanczos_nalpha", i64, [vp])
        sig(g + "lanczos_nbeta", i64, [vp])
        sig(g + "lanczos_get_alpha_beta", None, [vp, vp, vp])
        sig(g + "arnoldi_set_params", None, [vp, vp, dbl])
        sig(g + "arnoldi_residue", dbl, [vp])
        sig(g + "arnoldi_hess_size", i64, [vp])
        sig(g + "arnoldi_hessenberg", None, [vp, vp])
        sig(g + "arnoldi_matrix_cols", i64, [vp])
        s = g + "lsolver_"
        sig(s + "create", vp, [])
        sig(s + "destroy", None, [vp])
        sig(s + "set_op", None, [vp, vp, vp, i64])
        sig(s + "set_init", None, [vp, vp, i64])
        sig(s + "set_init_seeded", None, [vp, u32])
        sig(s + "get_init", None, [vp, vp])
        sig(s + "init_size", i64, [vp])
        sig(s + "clear_ortho", None, [vp])
        sig(s + "add_ortho", None, [vp, vp, i64])
        sig(s + "set_params", None, [vp, dbl, i64, dbl, dbl, i64, i64, i64, cint])
        sig(s + "set_indices", None, [vp, vp, i64])
        sig(s + "compute", cint, [vp])
        sig(s + "continue", cint, [vp])
        sig(s + "clear", None, [vp])
        sig(s + "iterations", i64, [vp])
        sig(s + "nvectors", i64, [vp])
        sig(s + "get_vector", None, [vp, i64, vp])
        sig(s + "nalpha", i64, [vp])
        sig(s + "nbeta", i64, [vp])
        sig(s + "get_alpha_beta", None, [vp, vp, vp])
        sig(s + "neig", i64, [vp])
        sig(s + "get_eigenvalues", None, [vp, vp])
        sig(s + "eigenvectors_shape", None, [vp, vp, vp])
        sig(s + "get_eigenvectors", None, [vp, vp])
        sig(s + "tri_size", i64, [vp])
        sig(s + "get_tri", None, [vp, vp, vp])
        sig(s + "nlog", i64, [vp])
        sig(s + "get_log", cint, [vp, i64, C.c_char_p, i64])
        sig(s + "has_warn", i64, [vp])
        sig(s + "has_error", i64, [vp])
        sig(s + "last_error", cint, [vp, C.c_char_p, i64])
        sig(s + "convlog_len", i64, [vp, i64])
        sig(s + "get_convlog", None, [vp, i64, vp])
        sig(g + "exp_with_eigens", None, [vp, vp, i64, vp, i64, i64, vp, vp])
        sig(g + "exp_with_lanczos", None, [vp, vp, vp])
        sig(g + "exp_taylor", None, [cint, vp, vp, vp, i64, dbl, vp, vp, dbl, i64])
    s = "ref_z_asolver_"
    sig(s + "create", vp, [])
    sig(s + "destroy", None, [vp])
    sig(s + "set_op", None, [vp, vp, vp, i64])
    sig(s + "set_init", None, [vp, vp, i64])
    sig(s + "get_init", None, [vp, vp])
    sig(s + "init_size", i64, [vp])
    sig(s + "clear_ortho", None, [vp])
    sig(s + "add_ortho", None, [vp, vp, i64])
    sig(s + "set_params", None, [vp, vp, dbl, dbl, i64, i64, i64, cint])
    sig(s + "set_indices", None, [vp, vp, i64])
    sig(s + "compute", cint, [vp])
    sig(s + "continue", cint, [vp])
    sig(s + "iterations", i64, [vp])
    sig(s + "nvectors", i64, [vp])
    sig(s + "get_vector", None, [vp, i64, vp])
    sig(s + "residue", dbl, [vp])
    sig(s + "hess_size", i64, [vp])
    sig(s + "hessenberg", None, [vp, vp])
    sig(s + "neig", i64, [vp])
    sig(s + "get_eigenvalues", None, [vp, vp])
    sig(s + "eigenvectors_shape", None, [vp, vp, vp])
    sig(s + "get_eigenvectors", None, [vp, vp])
    sig(s + "get_des", None, [vp, vp, vp])
    sig(s + "yh_rows", i64, [vp])
    sig(s + "get_yh", None, [vp, vp])
    sig(s + "convlog_len", i64, [vp, i64])
    sig(s + "get_convlog", None, [vp, i64, vp])
    sig(s + "nlog", i64, [vp])
    sig(s + "get_log", cint, [vp, i64, C.c_char_p, i64])
    sig(s + "has_warn", i64, [vp])
    sig(s + "last_error", cint, [vp, C.c_char_p, i64])


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(t):
    lib().ref_set_num_threads(int(t))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _op_callback(op):
    """(function pointer, user pointer) that make the reference's MatMulFunction call a core.Operator."""
    if op is None:
        return None, None
    fn = getattr(core.lib(), "orc_%s_op_apply" % op.p)
    return C.cast(fn, C.c_void_p), C.c_void_p(op.h)


def _string(getter, *args):
    need = getter(*args, None, 0)
    buf = C.create_string_buffer(need + 1)
    getter(*args, buf, need + 1)
    return buf.value.decode()


def default_vector(n, prefix="d"):
    """LanczosBase<Scalar>::setInitialVector() of the reference itself (lanczos.hpp:214-218)."""
    out = np.empty(n, dtype=_DT[prefix])
    getattr(lib(), "ref_%s_default_vector" % prefix)(n, _ptr(out))
    return out


def arnoldi_default_vector(n, prefix="z"):
    out = np.empty(n, dtype=_DT[prefix])
    getattr(lib(), "ref_%s_arnoldi_default_vector" % prefix)(n, _ptr(out))
    return out


def seeded_vector(seed, n, prefix="d"):
    """LanczosBase<Scalar>::makeRandomVector(std::mt19937(seed), n) (lanczos.hpp:124-135)."""
    out = np.empty(n, dtype=_DT[prefix])
    getattr(lib(), "ref_%s_seeded_vector" % prefix)(seed, n, _ptr(out))
    return out


class _Base:
    kind = None

    def __init__(self, prefix="d"):
        self.p = prefix
        self.dt = _DT[prefix]
        self._f = lambda name: getattr(lib(), "ref_%s_%s_%s" % (prefix, self.kind, name))
        self.h = self._f("create")()
        self.op = None
        self.n = 0

    def set_op(self, op, n=None):
        self.op = op
        self.n = int(op.n if (op is not None and n is None) else (n or 0))
        fn, user = _op_callback(op)
        self._f("set_op")(self.h, fn, user, self.n)

    def set_init(self, v):
        v = np.ascontiguousarray(v, dtype=self.dt)
        self._f("set_init")(self.h, _ptr(v), v.size)

    def add_ortho(self, v):
        v = np.ascontiguousarray(v, dtype=self.dt)
        self._f("add_ortho")(self.h, _ptr(v), v.size)

    def clear_steps(self):
        self._f("clear_steps")(self.h)

    def step(self):
        return bool(self._f("step")(self.h))

    def utmost(self):
        return bool(self._f("utmost")(self.h))

    @property
    def iterations(self):
        return int(self._f("iterations")(self.h))

    @property
    def nvectors(self):
        return int(self._f("nvectors")(self.h))

    def vector(self, k):
        out = np.empty(self.n, dtype=self.dt)
        self._f("get_vector")(self.h, k, _ptr(out))
        return out

    def __del__(self):
        try:
            if self.h:
                self._f("destroy")(self.h)
                self.h = None
        except Exception:
            pass


class LanczosBase(_Base):
    """cmpt::EigenEx::LanczosBase<Scalar> itself (lanczos.hpp:104-461)."""

    kind = "lanczos"

    def set_params(self, shift=0.0, interval=1, threshold=1e-12):
        self._f("set_params")(self.h, float(shift), int(interval), float(threshold))

    def alpha_beta(self):
        na, nb = int(self._f("nalpha")(self.h)), int(self._f("nbeta")(self.h))
        a, b = np.empty(na), np.empty(nb)
        self._f("get_alpha_beta")(self.h, _ptr(a), _ptr(b))
        return a, b


class ArnoldiBase(_Base):
    """cmpt::EigenEx::ArnoldiBase<Scalar> itself (arnoldi.hpp:53-438)."""

    kind = "arnoldi"

    def set_params(self, shift=0.0, threshold=1e-12):
        s = np.array([shift], dtype=self.dt)
        self._f("set_params")(self.h, _ptr(s), float(threshold))

    @property
    def residue(self):
        return float(self._f("residue")(self.h))

    def hessenberg(self):
        hs = int(self._f("hess_size")(self.h))
        out = np.zeros((hs, hs), dtype=self.dt, order="F")
        if hs:
            self._f("hessenberg")(self.h, _ptr(out))
        return out


class _Solver:
    """Settings shared by the two solver wrappers; attribute names as in oracle/reference_solvers.py."""

    unlimited = -1

    def _defaults(self):
        self.min_iterations = 1
        self.max_iterations = -1
        self.tolerance = 1e-12
        self.indices_for_convergence = [0]
        self.max_eigenvalues = -1
        self.compute_eigenvectors_on = True
        self.shift = 0.0
        self.interval = 1
        self.threshold = 1e-12
        self.op = None
        self.n = 0
        self.init = None
        self.ortho = []

    def set_matrix_multiplication(self, op, n=None):
        self.op = op
        self.n = int(op.n if (op is not None and n is None) else (n or 0))

    def _push_common(self):
        fn, user = _op_callback(self.op)
        self._f("set_op")(self.h, fn, user, self.n)
        if self.init is not None:
            v = np.ascontiguousarray(self.init, dtype=self.dt)
            self._f("set_init")(self.h, _ptr(v), v.size)
        self._f("clear_ortho")(self.h)
        for o in self.ortho:
            o = np.ascontiguousarray(o, dtype=self.dt)
            self._f("add_ortho")(self.h, _ptr(o), o.size)
        idx = np.asarray(self.indices_for_convergence, dtype=np.int64)
        self._f("set_indices")(self.h, _ptr(idx), idx.size)

    def _run(self, what):
        self._push()
        rc = self._f(what)(self.h)
        if rc != 0:
            raise RuntimeError("reference threw: " + _string(self._f("last_error"), self.h))
        self._pull()
        return 0

    def compute(self):
        return self._run("compute")

    def continue_to_compute(self):
        return self._run("continue")

    @property
    def iterations(self):
        return int(self._f("iterations")(self.h))

    @property
    def nvectors(self):
        return int(self._f("nvectors")(self.h))

    def vector(self, k):
        out = np.empty(self.n, dtype=self.dt)
        self._f("get_vector")(self.h, k, _ptr(out))
        return out

    def _pull_log(self):
        self.log = [_string(self._f("get_log"), self.h, i) for i in range(int(self._f("nlog")(self.h)))]

    def has_warn(self):
        return int(self._f("has_warn")(self.h))

    def __del__(self):
        try:
            if self.h:
                self._f("destroy")(self.h)
                self.h = None
        except Exception:
            pass


class LanczosEigenSolver(_Solver):
    """cmpt::EigenEx::LanczosEigenSolver<Scalar> itself (lanczos.hpp:468-927)."""

    def __init__(self, prefix="d"):
        self.p = prefix
        self.dt = _DT[prefix]
        self._f = lambda name: getattr(lib(), "ref_%s_lsolver_%s" % (prefix, name))
        self.h = self._f("create")()
        self._defaults()
        self.eigenvalues = np.zeros(0)
        self.eigenvectors = np.zeros((0, 0))
        self.log = []
        self.convergence_log = {}

    def set_init_seeded(self, seed):
        """setInitialVector(lanczosBase().makeRandomVector(std::mt19937(seed), n)) as sample_lanczos2.cpp:39,53."""
        fn, user = _op_callback(self.op)
        self._f("set_op")(self.h, fn, user, self.n)
        self._f("set_init_seeded")(self.h, seed)
        self.init = self._get_init()

    def _get_init(self):
        out = np.empty(int(self._f("init_size")(self.h)), dtype=self.dt)
        self._f("get_init")(self.h, _ptr(out))
        return out

    def _push(self):
        self._push_common()
        self._f("set_params")(self.h, float(self.shift), int(self.interval), float(self.threshold), float(self.tolerance),
                              int(self.min_iterations), int(self.max_iterations), int(self.max_eigenvalues),
                              int(bool(self.compute_eigenvectors_on)))

    def _pull(self):
        self.init = self._get_init()
        nev = int(self._f("neig")(self.h))
        self.eigenvalues = np.empty(nev)
        self._f("get_eigenvalues")(self.h, _ptr(self.eigenvalues))
        r, c = C.c_int64(), C.c_int64()
        self._f("eigenvectors_shape")(self.h, C.byref(r), C.byref(c))
        self.eigenvectors = np.empty((r.value, c.value), dtype=self.dt, order="F")
        if r.value * c.value:
            self._f("get_eigenvectors")(self.h, _ptr(self.eigenvectors))
        self._pull_log()
        self.convergence_log = {}
        for idx in self.indices_for_convergence:
            ln = int(self._f("convlog_len")(self.h, int(idx)))
            if ln >= 0:
                a = np.empty(ln)
                self._f("get_convlog")(self.h, int(idx), _ptr(a))
                self.convergence_log[idx] = list(a)

    def alpha_beta(self):
        na, nb = int(self._f("nalpha")(self.h)), int(self._f("nbeta")(self.h))
        a, b = np.empty(na), np.empty(nb)
        self._f("get_alpha_beta")(self.h, _ptr(a), _ptr(b))
        return a, b

    def tridiagonal_eigensystem(self):
        """es_tri(): Ritz values and the eigenvectors S of T_k after the last trip."""
        k = int(self._f("tri_size")(self.h))
        theta = np.empty(k)
        S = np.empty((k, k), dtype=self.dt, order="F")
        self._f("get_tri")(self.h, _ptr(theta), _ptr(S) if k else None)
        return theta, S

    def has_error(self):
        return int(self._f("has_error")(self.h))

    def exp_with_lanczos(self, x):
        """LanczosExponentialSolver<Scalar>::solveWithLanczos(x, *this, out) (lanczos.hpp:1061-1075)."""
        self._push()
        xs = np.array([x], dtype=self.dt)
        out = np.empty(self.n, dtype=self.dt)
        getattr(lib(), "ref_%s_exp_with_lanczos" % self.p)(_ptr(xs), self.h, _ptr(out))
        self._pull()
        return out


def exp_with_eigens(x, eivals, eivecs, max_expand, vin, prefix="d"):
    """LanczosExponentialSolver<Scalar>::solveWithEigens (lanczos.hpp:1024-1054)."""
    dt = _DT[prefix]
    xs = np.array([x], dtype=dt)
    w = np.ascontiguousarray(eivals, dtype=np.float64)
    y = np.asfortranarray(eivecs, dtype=dt)
    vin = np.ascontiguousarray(vin, dtype=dt)
    out = np.empty(vin.size, dtype=dt)
    getattr(lib(), "ref_%s_exp_with_eigens" % prefix)(_ptr(xs), _ptr(w), w.size, _ptr(y), vin.size, int(max_expand), _ptr(vin), _ptr(out))
    return out


def exp_taylor(x, op, radius, vin, error=1e-14, max_expansion=-1, auto_division=False):
    """solveWithTaylorNoDivision / solveWithTaylorAutoDivision (lanczos.hpp:1085-1161)."""
    dt = _DT[op.p]
    xs = np.array([x], dtype=dt)
    vin = np.ascontiguousarray(vin, dtype=dt)
    out = np.empty(vin.size, dtype=dt)
    fn, user = _op_callback(op)
    getattr(lib(), "ref_%s_exp_taylor" % op.p)(int(bool(auto_division)), _ptr(xs), fn, user, op.n, float(radius), _ptr(vin), _ptr(out),
                                               float(error), int(max_expansion))
    return out


class ArnoldiEigenSolver(_Solver):
    """cmpt::EigenEx::ArnoldiEigenSolver<std::complex<double>> itself (arnoldi.hpp:444-1027) — the only Scalar
    the reference's class compiles for (arnoldi.hpp:857,864)."""

    def __init__(self, prefix="z"):
        if prefix != "z":
            raise ValueError("the reference's ArnoldiEigenSolver only compiles for complex Scalar")
        self.p = prefix
        self.dt = np.complex128
        self._f = lambda name: getattr(lib(), "ref_z_asolver_%s" % name)
        self.h = self._f("create")()
        self._defaults()
        self.eigenvalues = np.zeros(0, complex)
        self.eigenvectors = np.zeros((0, 0), complex)
        self.eigenvectors_h = np.zeros((0, 0), complex)
        self.hessenberg = np.zeros((0, 0), complex)
        self.log = []
        self.convergence_log = {}

    def _push(self):
        self._push_common()
        s = np.array([self.shift], dtype=np.complex128)
        self._f("set_params")(self.h, _ptr(s), float(self.threshold), float(self.tolerance), int(self.min_iterations),
                              int(self.max_iterations), int(self.max_eigenvalues), int(bool(self.compute_eigenvectors_on)))

    def _pull(self):
        out = np.empty(int(self._f("init_size")(self.h)), dtype=np.complex128)
        self._f("get_init")(self.h, _ptr(out))
        self.init = out
        nev = int(self._f("neig")(self.h))
        self.eigenvalues = np.empty(nev, complex)
        self._f("get_eigenvalues")(self.h, _ptr(self.eigenvalues))
        r, c = C.c_int64(), C.c_int64()
        self._f("eigenvectors_shape")(self.h, C.byref(r), C.byref(c))
        self.eigenvectors = np.empty((r.value, c.value), dtype=complex, order="F")
        if r.value * c.value:
            self._f("get_eigenvectors")(self.h, _ptr(self.eigenvectors))
        hs = int(self._f("hess_size")(self.h))
        self.hessenberg = np.zeros((hs, hs), complex, order="F")
        if hs:
            self._f("hessenberg")(self.h, _ptr(self.hessenberg))
        k = int(self._f("yh_rows")(self.h))
        self.eigenvectors_h = np.zeros((k, k), complex, order="F")
        if k:
            self._f("get_yh")(self.h, _ptr(self.eigenvectors_h))
        self._pull_log()
        self.convergence_log = {}
        for idx in self.indices_for_convergence:
            ln = int(self._f("convlog_len")(self.h, int(idx)))
            if ln >= 0:
                a = np.empty(ln, complex)
                self._f("get_convlog")(self.h, int(idx), _ptr(a))
                self.convergence_log[idx] = list(a)

    @property
    def residue(self):
        return float(self._f("residue")(self.h))
