"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY).

The library is the CPU restatement of the reference's Krylov step functions
(see krylov_oracle.cpp for the file:line map).  Importers: tests/, the smoke
check in __graft_entry__.py and the cpu_baseline / --impl reference legs of
bench.py.  The product (cmpt_eigenex_b200/, include/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_DT = {"d": np.float64, "z": np.complex128}


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++ only, no reference sources)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "krylov_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], env=env)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _declare(_LIB)
    return _LIB


def _declare(L):
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]
    for p in "dz":
        g = lambda name: getattr(L, "orc_%s_%s" % (p, name))
        g("op_csr").restype = vp
        g("op_csr").argtypes = [i64, vp, vp, vp]
        g("op_dense").restype = vp
        g("op_dense").argtypes = [i64, vp]
        g("op_heisenberg").restype = vp
        g("op_heisenberg").argtypes = [C.c_int, dbl, C.c_int]
        g("op_callback").restype = vp
        g("op_callback").argtypes = [vp, vp]
        g("op_apply").argtypes = [vp, vp, vp]
        g("op_destroy").argtypes = [vp]
        g("default_vector").argtypes = [i64, vp]
        g("seeded_vector").argtypes = [C.c_uint32, i64, vp]
        for k in ("lanczos", "arnoldi"):
            g(k + "_create").restype = vp
            g(k + "_destroy").argtypes = [vp]
            g(k + "_set_op").argtypes = [vp, vp, i64]
            g(k + "_set_init").argtypes = [vp, vp, i64]
            g(k + "_add_ortho").argtypes = [vp, vp, i64]
            g(k + "_clear_steps").argtypes = [vp]
            g(k + "_step").argtypes = [vp]
            g(k + "_step").restype = C.c_int
            g(k + "_utmost").argtypes = [vp]
            g(k + "_utmost").restype = C.c_int
            g(k + "_iterations").argtypes = [vp]
            g(k + "_iterations").restype = i64
            g(k + "_nvectors").argtypes = [vp]
            g(k + "_nvectors").restype = i64
            g(k + "_get_vector").argtypes = [vp, i64, vp]
            g(k + "_assemble").argtypes = [vp, vp, i64, i64, i64, vp]
        g("lanczos_set_params").argtypes = [vp, dbl, i64, dbl]
        g("lanczos_nalpha").argtypes = [vp]
        g("lanczos_nalpha").restype = i64
        g("lanczos_nbeta").argtypes = [vp]
        g("lanczos_nbeta").restype = i64
        g("lanczos_get_alpha_beta").argtypes = [vp, vp, vp]
        g("arnoldi_set_params").argtypes = [vp, vp, dbl]
        g("arnoldi_residue").argtypes = [vp]
        g("arnoldi_residue").restype = dbl
        g("arnoldi_hess_size").argtypes = [vp]
        g("arnoldi_hess_size").restype = i64
        g("arnoldi_hessenberg").argtypes = [vp, vp]


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(t):
    lib().orc_set_num_threads(int(t))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Operator:
    """Host operator handle: CSR, dense (row-major), Heisenberg chain or a Python callback."""

    def __init__(self, prefix, handle, n, keep=None):
        self.p, self.h, self.n, self._keep = prefix, handle, int(n), keep

    @staticmethod
    def csr(rowptr, col, val):
        val = np.ascontiguousarray(val)
        p = "z" if np.iscomplexobj(val) else "d"
        val = val.astype(_DT[p], copy=False)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        n = rowptr.size - 1
        h = getattr(lib(), "orc_%s_op_csr" % p)(n, _ptr(rowptr), _ptr(col), _ptr(val))
        return Operator(p, h, n)

    @staticmethod
    def dense(a):
        a = np.asarray(a)
        p = "z" if np.iscomplexobj(a) else "d"
        a = np.ascontiguousarray(a, dtype=_DT[p])
        h = getattr(lib(), "orc_%s_op_dense" % p)(a.shape[0], _ptr(a))
        return Operator(p, h, a.shape[0])

    @staticmethod
    def heisenberg(L, J=1.0, pbc=True, prefix="d"):
        h = getattr(lib(), "orc_%s_op_heisenberg" % prefix)(int(L), float(J), int(bool(pbc)))
        return Operator(prefix, h, 1 << L)

    @staticmethod
    def callback(fn, n, prefix="d"):
        dt = _DT[prefix]

        def tramp(pin, pout, _user):
            x = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_double)), shape=(n * (2 if prefix == "z" else 1),)).view(dt)
            y = np.ctypeslib.as_array(C.cast(pout, C.POINTER(C.c_double)), shape=(n * (2 if prefix == "z" else 1),)).view(dt)
            y[:] = fn(x)

        cb = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)(tramp)
        h = getattr(lib(), "orc_%s_op_callback" % prefix)(C.cast(cb, C.c_void_p), None)
        return Operator(prefix, h, n, keep=cb)

    def apply(self, x):
        x = np.ascontiguousarray(x, dtype=_DT[self.p])
        y = np.empty_like(x)
        getattr(lib(), "orc_%s_op_apply" % self.p)(self.h, _ptr(x), _ptr(y))
        return y

    def __del__(self):
        try:
            if self.h:
                getattr(lib(), "orc_%s_op_destroy" % self.p)(self.h)
                self.h = None
        except Exception:
            pass


def default_vector(n, prefix="d"):
    """std::mt19937 default seed + std::normal_distribution, normalised (lanczos.hpp:214-218)."""
    out = np.empty(n, dtype=_DT[prefix])
    getattr(lib(), "orc_%s_default_vector" % prefix)(n, _ptr(out))
    return out


def seeded_vector(seed, n, prefix="d"):
    """makeRandomVector(std::mt19937(seed), n) (lanczos.hpp:124-135, sample_lanczos2.cpp:39,53)."""
    out = np.empty(n, dtype=_DT[prefix])
    getattr(lib(), "orc_%s_seeded_vector" % prefix)(seed, n, _ptr(out))
    return out


class _Base:
    kind = None

    def __init__(self, prefix="d"):
        self.p = prefix
        self.dt = _DT[prefix]
        self._f = lambda name: getattr(lib(), "orc_%s_%s_%s" % (prefix, self.kind, name))
        self.h = self._f("create")()
        self.op = None
        self.n = 0

    def set_op(self, op, n=None):
        self.op = op
        self.n = int(op.n if (op is not None and n is None) else (n or 0))
        self._f("set_op")(self.h, op.h if op is not None else None, self.n)

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
    """LanczosBase<Scalar> state machine (lanczos.hpp:104-461)."""

    kind = "lanczos"

    def set_params(self, shift=0.0, interval=1, threshold=1e-12):
        self._f("set_params")(self.h, float(shift), int(interval), float(threshold))

    def alpha_beta(self):
        na, nb = int(self._f("nalpha")(self.h)), int(self._f("nbeta")(self.h))
        a, b = np.empty(na), np.empty(nb)
        self._f("get_alpha_beta")(self.h, _ptr(a), _ptr(b))
        return a, b

    def assemble(self, coef, nev):
        coef = np.asfortranarray(coef, dtype=self.dt)
        out = np.empty((self.n, nev), dtype=self.dt, order="F")
        self._f("assemble")(self.h, _ptr(coef), coef.shape[0], coef.shape[0], nev, _ptr(out))
        return out


class ArnoldiBase(_Base):
    """ArnoldiBase<Scalar> state machine (arnoldi.hpp:53-438)."""

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

    def assemble(self, coef, nev):
        coef = np.asfortranarray(coef, dtype=np.complex128)
        out = np.empty((self.n, nev), dtype=np.complex128, order="F")
        self._f("assemble")(self.h, _ptr(coef), coef.shape[0], coef.shape[0], nev, _ptr(out))
        return out
