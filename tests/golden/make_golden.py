"""Generates tests/golden/kat.json and tests/golden/oracle_traces.npz.

The reference ships no golden vectors and cannot be built here (it needs Eigen3), so two kinds of fixtures are
committed instead:

* kat.json — known answers that depend on neither the oracle nor the product: closed-form spectra and the
  eigen-decompositions of the reference's sample matrices computed with LAPACK (numpy), the Heisenberg-ring ground
  state energies of SURVEY.md Appendix E (ARPACK, tol 1e-13) recomputed here with scipy for L <= 16.
* oracle_traces.npz — alpha/beta, Hessenberg columns, Ritz values and convergence logs of the CPU oracle
  (oracle/, the restatement of the reference's algorithm) on small seeded problems.  They pin the oracle against
  silent changes and give the GPU tests a fixed target that does not require running the oracle.

usage: python tests/golden/make_golden.py     (CPU only; needs oracle/liboracle.so built)
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from cmpt_eigenex_b200 import synthetic as syn  # noqa: E402
from oracle import core  # noqa: E402
from oracle import reference_solvers as rs  # noqa: E402


def csr_to_scipy(rp, c, v, n):
    return sp.csr_matrix((v, c, rp), shape=(n, n))


def main():
    kat = {}
    # sample_lanczos1.cpp:13-17
    H = np.array([[1.0, 0.5, 0.0], [0.5, 2.0, 0.5], [0.0, 0.5, 3.0]])
    w, X = np.linalg.eigh(H)
    X = X * np.sign(X[0])  # phase convention of lanczos.hpp:806-813: first component positive
    kat["sample_lanczos1"] = {"matrix": H.tolist(), "eigenvalues": w.tolist(), "eigenvectors_columns": X.T.tolist(),
                              "closed_form": [2 - 1.5 ** 0.5, 2.0, 2 + 1.5 ** 0.5]}
    # sample_lanczos2.cpp:19-34: spectrum 2 cos(k pi / 201)
    n = 200
    kat["sample_lanczos2"] = {"n": n, "lowest_ten": sorted(2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1)))[:10]}
    # 2D Dirichlet Laplacian, 3D convection-diffusion (closed forms of SURVEY.md §8(c))
    kat["laplacian2d"] = {str(N): syn.laplacian2d_eigenvalues(N, 6).tolist() for N in (12, 40, 64, 4096)}
    kat["convdiff3d_leading"] = {str(M): float(syn.convdiff3d_eigenvalues(M, count=1)[0]) for M in (5, 9, 24, 256)}
    # Heisenberg ring ground states: table of SURVEY.md Appendix E, re-derived with ARPACK where cheap
    e0 = {}
    for L in (8, 12, 16):
        rp, c, v = syn.heisenberg_csr(L)
        e0[str(L)] = float(spla.eigsh(csr_to_scipy(rp, c, v, 1 << L), k=1, which="SA", tol=1e-13)[0][0])
    kat["heisenberg_ring_E0_arpack"] = e0
    kat["heisenberg_ring_E0_survey"] = {str(k): v for k, v in syn.HEISENBERG_RING_E0.items()}
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---- oracle traces ----
    core.set_num_threads(1)
    tr = {}
    # Lanczos, real: 2D Laplacian 12 x 12, m = 30, seed 7
    N, m = 12, 30
    rp, c, v = syn.laplacian2d_csr(N)
    es = rs.LanczosEigenSolver("d")
    es.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    es.init = syn.start_vector(N * N, seed=7)
    es.min_iterations = es.max_iterations = m
    es.max_eigenvalues = 4
    es.indices_for_convergence = [0, 1]
    es.compute()
    a, b = es.alpha_beta()
    tr["lap12_alpha"], tr["lap12_beta"], tr["lap12_eigenvalues"] = a, b, es.eigenvalues
    tr["lap12_convlog0"] = np.array(es.convergence_log[0])
    tr["lap12_eigenvectors"] = es.eigenvectors
    # Lanczos, complex: Hermitian chain n = 60, m = 25, libstdc++ mt19937(1) start vector as in sample_lanczos2.cpp
    rp, c, v = syn.hermitian_chain_csr(60)
    es = rs.LanczosEigenSolver("z")
    es.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    es.init = core.seeded_vector(1, 60, "z")
    es.min_iterations = es.max_iterations = 25
    es.max_eigenvalues = 3
    es.compute()
    a, b = es.alpha_beta()
    tr["chain60_init"], tr["chain60_alpha"], tr["chain60_beta"], tr["chain60_eigenvalues"] = es.init, a, b, es.eigenvalues
    # Arnoldi, real operator: convection-diffusion 5^3, m = 12, seed 7
    M, m = 5, 12
    rp, c, v = syn.convdiff3d_csr(M)
    es = rs.ArnoldiEigenSolver("d")
    es.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    es.init = syn.start_vector(M ** 3, seed=7)
    es.min_iterations = es.max_iterations = m
    es.max_eigenvalues = 3
    es.compute()
    tr["cd5_hessenberg"], tr["cd5_eigenvalues"] = es.hessenberg, es.eigenvalues
    np.savez(os.path.join(HERE, "oracle_traces.npz"), **tr)
    print("wrote kat.json and oracle_traces.npz:", {k: np.asarray(v).shape for k, v in tr.items()})


if __name__ == "__main__":
    main()
