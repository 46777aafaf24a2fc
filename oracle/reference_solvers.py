"""Driver loops of the reference's solver classes, restated (TEST INFRASTRUCTURE ONLY).

LanczosEigenSolver  <- lanczos.hpp:468-927   (compute :717-736, continueToCompute :701-712,
                       mainCalculation_ :740-823, getFormalIndex :837-847,
                       updateConvergenceLog_ :853-864, isConverged_ :869-896)
ArnoldiEigenSolver  <- arnoldi.hpp:444-1027  (compute :741-760, continueToCompute :725-736,
                       mainCalculation_ :764-873, sort :813-822, isConverged_ :969-996)

The Krylov steps run in liboracle.so (oracle/core.py).  The m x m Ritz problems, which
the reference hands to Eigen (SelfAdjointEigenSolver::computeFromTridiagonal,
ComplexEigenSolver / EigenSolver; Eigen is a third-party dependency that is not in
/root/reference and is un-pinned there), are handed to LAPACK through scipy here.  Any
backward-stable dense solver agrees with Eigen's to ~eps*||T||, far inside the 1e-10 gate.

parity unpinned: the reference ships no golden vectors; these classes
are pinned on analytic known answers in tests/test_oracle_kat.py.
"""
import numpy as np
import scipy.linalg as sla

from . import core

UNLIMITED = -1
HEAD_ERROR = "ERROR     "
HEAD_WARN = "WARN      "
HEAD_INFO = "INFO      "
HEAD_DEBUG = "DEBUG     "


def formal_index(i, n):
    """lanczos.hpp:837-847 / arnoldi.hpp:938-948."""
    if -n <= i < 0:
        return n - (-i - 1) % n - 1
    if 0 <= i < n:
        return i % n
    return -1


def eig_tridiagonal(alpha, beta):
    """Stand-in for computeFromTridiagonal(alpha, beta) (lanczos.hpp:779-781): ascending
    eigenvalues and eigenvectors of T_k; only the first k-1 sub-diagonal entries are read."""
    k = len(alpha)
    if k == 0:
        return np.zeros(0), np.zeros((0, 0))
    if k == 1:
        return np.array([alpha[0]], dtype=float), np.ones((1, 1))
    w, v = sla.eigh_tridiagonal(np.asarray(alpha, float), np.asarray(beta[: k - 1], float))
    return w, v


class LanczosEigenSolver:
    unlimited = UNLIMITED

    def __init__(self, prefix="d"):
        self.p = prefix
        self.base = core.LanczosBase(prefix)
        self.set_all_settings_default()
        self.eigenvalues = np.zeros(0)
        self.eigenvectors = np.zeros((0, 0))
        self.log = []
        self.convergence_log = {}
        self._theta = np.zeros(0)
        self._S = np.zeros((0, 0))

    # lanczos.hpp:657-668 and :260-271
    def set_all_settings_default(self):
        self.min_iterations = 1
        self.max_iterations = UNLIMITED
        self.tolerance = 1e-12
        self.indices_for_convergence = [0]
        self.max_eigenvalues = UNLIMITED
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
        self.n = op.n if n is None else n

    def _push_settings(self):
        b = self.base
        b.set_op(self.op, self.n)
        b.set_params(self.shift, self.interval, self.threshold)

    def clear_computed_data(self):  # lanczos.hpp:675-682
        self.base.clear_steps()
        self.eigenvalues = np.zeros(0)
        self.eigenvectors = np.zeros((0, 0))
        self.log = []
        self.convergence_log = {}

    def compute(self):  # lanczos.hpp:717-736
        self.log.append(HEAD_INFO + "EigenSolver<ScalarType>::compute(...) was called")
        self.clear_computed_data()
        # A fresh handle keeps deflation vectors / start vector in sync with the settings.
        self.base = core.LanczosBase(self.p)
        self._push_settings()
        for o in self.ortho:
            self.base.add_ortho(o)
        if self.init is None or len(self.init) != self.n:
            self.log.append(HEAD_INFO + "in compute(), initial_vector is empty or invalid, then set at random")
            self.init = core.default_vector(self.n, self.p) if self.n > 0 else np.zeros(0)
        self.base.set_init(self.init)
        ret = self._main()
        self.log.append(HEAD_INFO + "EigenSolver<ScalarType>::compute(...) finish computing")
        return ret

    def continue_to_compute(self):  # lanczos.hpp:701-712
        self.log.append(HEAD_INFO + "EigenSolver<ScalarType>::continueToCompute(...) was called")
        if self.base.nvectors == 0:
            return self.compute()
        self._push_settings()
        ret = self._main()
        self.log.append(HEAD_INFO + "EigenSolver<ScalarType>::compute(...) finish computing")
        return ret

    @property
    def iterations(self):
        return self.base.iterations

    def alpha_beta(self):
        return self.base.alpha_beta()

    def _update_convergence_log(self):  # lanczos.hpp:853-864
        for idx in self.indices_for_convergence:
            i = formal_index(idx, len(self._theta))
            if i < 0:
                continue
            self.convergence_log.setdefault(idx, []).append(float(self._theta[i]))

    def _is_converged(self):  # lanczos.hpp:869-896
        if len(self._theta) < 2:
            return False
        scale = self._theta[0] - self._theta[-1]
        for idx in self.indices_for_convergence:
            edge = self.convergence_log.get(idx)
            if edge is None or len(edge) < 2:
                return False
            if abs((edge[-1] - edge[-2]) / scale) > self.tolerance:
                return False
        return True

    def _main(self):  # lanczos.hpp:740-823
        b = self.base
        self._theta, self._S = eig_tridiagonal([], [])
        init_fail = False
        while True:
            self._update_convergence_log()
            if init_fail:
                self.log.append(HEAD_INFO + "initial lanczosvector generation fail")
                break
            if b.utmost():
                self.log.append(HEAD_INFO + "lanczos steps finished with threshold")
                self.log.append(HEAD_INFO + "lanczos steps achieved full of Krylov subspace")
                break
            if b.iterations >= self.min_iterations:
                if b.iterations == self.max_iterations:
                    self.log.append(HEAD_WARN + "lanczos steps achieved maxIterations")
                    break
                if self._is_converged():
                    self.log.append(HEAD_INFO + "lanczos steps converged with tolerance")
                    break
            b.step()
            if b.nvectors == 0:
                init_fail = True
            a, be = b.alpha_beta()
            self._theta, self._S = eig_tridiagonal(a, be)
        nev = len(self._theta)
        if self.max_eigenvalues != UNLIMITED and self.max_eigenvalues < nev:
            nev = self.max_eigenvalues
        self.eigenvalues = np.array(self._theta[:nev]) - self.shift
        if self.compute_eigenvectors_on:
            if nev > 0:
                self.eigenvectors = b.assemble(self._S[:, :nev], nev)
            else:
                self.eigenvectors = np.zeros((self.n, 0), dtype=b.dt)
        else:
            self.eigenvectors = np.zeros((0, 0))
        return 0

    def ritz_residuals(self):
        """|beta_next * S(last, i)| — not in the reference; derivable from its exposed state:
        beta_next = ||A u_k - alpha_k u_k - beta_{k-1} u_{k-1}|| (the kept beta after a breakdown)."""
        a, be = self.base.alpha_beta()
        k = len(a)
        if k == 0:
            return np.zeros(0)
        if len(be) >= k:
            bl = be[k - 1]
        else:
            u = self.base.vector(k - 1)
            r = self.op.apply(u) + self.shift * u - a[k - 1] * u
            if k > 1:
                r = r - be[k - 2] * self.base.vector(k - 2)
            bl = np.linalg.norm(r)
        return np.abs(bl * self._S[k - 1, : len(self.eigenvalues)])

    def has_warn(self):
        return sum(1 for s in self.log if s.startswith(HEAD_WARN))

    def has_error(self):
        return sum(1 for s in self.log if s.startswith(HEAD_ERROR))


def exp_solve_with_eigens(x, eivals, eivecs, max_expand, vin):
    """LanczosExponentialSolver::solveWithEigens (lanczos.hpp:1024-1054): sum of exp(x E_n) <y_n|in> y_n over the
    first max_expand pairs, smallest weight first."""
    mx = min(max_expand, len(eivals), eivecs.shape[1])
    out = np.zeros(len(vin), dtype=np.result_type(eivecs.dtype, np.asarray(vin).dtype, type(x)))
    for n_ in range(mx):
        n = mx - n_ - 1 if np.real(x) < 0.0 else n_
        inner = np.vdot(eivecs[:, n], vin)
        out = out + np.exp(x * eivals[n]) * inner * eivecs[:, n]
    return out


def exp_solve_with_lanczos(x, es):
    """LanczosExponentialSolver::solveWithLanczos (lanczos.hpp:1061-1075)."""
    es.compute()
    return exp_solve_with_eigens(x, es.eigenvalues, es.eigenvectors, len(es.eigenvalues), es.init)


class ArnoldiEigenSolver:
    unlimited = UNLIMITED

    def __init__(self, prefix="z"):
        self.p = prefix
        self.base = core.ArnoldiBase(prefix)
        self.set_all_settings_default()
        self.eigenvalues = np.zeros(0, complex)
        self.eigenvectors = np.zeros((0, 0), complex)
        self.eigenvectors_h = np.zeros((0, 0), complex)
        self.hessenberg = np.zeros((0, 0))
        self.log = []
        self.convergence_log = {}

    def set_all_settings_default(self):  # arnoldi.hpp:681-692, :208-218
        self.min_iterations = 1
        self.max_iterations = UNLIMITED
        self.tolerance = 1e-12
        self.indices_for_convergence = [0]
        self.max_eigenvalues = UNLIMITED
        self.compute_eigenvectors_on = True
        self.shift = 0.0
        self.threshold = 1e-12
        self.op = None
        self.n = 0
        self.init = None
        self.ortho = []

    def set_matrix_multiplication(self, op, n=None):
        self.op = op
        self.n = op.n if n is None else n

    def _push_settings(self):
        self.base.set_op(self.op, self.n)
        self.base.set_params(self.shift, self.threshold)

    def clear_computed_data(self):  # arnoldi.hpp:699-706
        self.base.clear_steps()
        self.eigenvalues = np.zeros(0, complex)
        self.eigenvectors = np.zeros((0, 0), complex)
        self.log = []
        self.convergence_log = {}

    def compute(self):  # arnoldi.hpp:741-760
        self.log.append(HEAD_INFO + "ArnoldiEigenSolver<ScalarType>::compute(...) was called")
        self.clear_computed_data()
        self.base = core.ArnoldiBase(self.p)
        self._push_settings()
        for o in self.ortho:
            self.base.add_ortho(o)
        if self.init is None or len(self.init) != self.n:
            self.log.append(HEAD_INFO + "in compute(), initial_vector is empty or invalid, then set at random")
            self.init = core.default_vector(self.n, self.p) if self.n > 0 else np.zeros(0)
        self.base.set_init(self.init)
        ret = self._main()
        self.log.append(HEAD_INFO + "ArnoldiEigenSolver<ScalarType>::compute(...) finish computing")
        return ret

    def continue_to_compute(self):  # arnoldi.hpp:725-736
        self.log.append(HEAD_INFO + "ArnoldiEigenSolver<ScalarType>::continueToCompute(...) was called")
        if self.base.nvectors == 0:
            return self.compute()
        self._push_settings()
        ret = self._main()
        self.log.append(HEAD_INFO + "ArnoldiEigenSolver<ScalarType>::compute(...) finish computing")
        return ret

    @property
    def iterations(self):
        return self.base.iterations

    def _update_convergence_log(self):  # arnoldi.hpp:954-964
        for idx in self.indices_for_convergence:
            i = formal_index(idx, len(self.eigenvalues))
            if i < 0:
                continue
            self.convergence_log.setdefault(idx, []).append(complex(self.eigenvalues[i]))

    def _is_converged(self):  # arnoldi.hpp:969-996
        if len(self.eigenvalues) < 2:
            return False
        scale = abs(self.eigenvalues[0] - self.eigenvalues[-1])
        for idx in self.indices_for_convergence:
            edge = self.convergence_log.get(idx)
            if edge is None or len(edge) < 2:
                return False
            if abs((edge[-1] - edge[-2]) / scale) > self.tolerance:
                return False
        return True

    def _main(self):  # arnoldi.hpp:764-873
        b = self.base
        init_fail = False
        while True:
            self._update_convergence_log()
            if init_fail:
                self.log.append(HEAD_INFO + "initial arnoldivector generation fail")
                break
            if b.utmost():
                self.log.append(HEAD_INFO + "arnoldi steps finished with threshold")
                self.log.append(HEAD_INFO + "arnoldi steps achieved full of Krylov subspace")
                break
            if b.iterations >= self.min_iterations:
                if b.iterations == self.max_iterations:
                    self.log.append(HEAD_WARN + "arnoldi steps achieved maxIterations")
                    break
                if self._is_converged():
                    self.log.append(HEAD_INFO + "arnoldi steps converged with tolerance")
                    break
            b.step()
            if b.nvectors == 0:
                init_fail = True
            H = b.hessenberg()
            self.hessenberg = H
            if H.shape[0] == 0:
                self.eigenvalues = np.zeros(0, complex)
                self.eigenvectors_h = np.zeros((0, 0), complex)
            else:
                w, y = sla.eig(H)  # unit 2-norm columns, as Eigen's ComplexEigenSolver
                order = sorted(range(len(w)), key=lambda i: -abs(w[i]))  # :813-822, descending |lambda|
                self.eigenvalues = np.asarray(w, complex)[order]
                self.eigenvectors_h = np.asarray(y, complex)[:, order]
        nev = len(self.eigenvalues)
        if self.max_eigenvalues != UNLIMITED and self.max_eigenvalues < nev:
            nev = self.max_eigenvalues
        self.eigenvalues = self.eigenvalues[:nev] - self.shift
        if self.compute_eigenvectors_on:
            if nev > 0:
                self.eigenvectors = b.assemble(self.eigenvectors_h[:, :nev], nev)
            else:
                self.eigenvectors = np.zeros((self.n, 0), complex)
        else:
            self.eigenvectors = np.zeros((0, 0), complex)
        return 0

    def ritz_residuals(self):
        """residue * |Y(last, i)| — not in the reference; derivable from its exposed state."""
        if self.eigenvectors_h.shape[0] == 0:
            return np.zeros(0)
        return self.base.residue * np.abs(self.eigenvectors_h[-1, : len(self.eigenvalues)])

    def has_warn(self):
        return sum(1 for s in self.log if s.startswith(HEAD_WARN))
