"""Records the reference ITSELF on BASELINE.json's configs at FULL size -> tests/golden/ref_full_cfg<N>.npz.

oracle/_ref/libref.so (the unmodified reference classes, see oracle/Makefile) is run on the host on
  cfg 1  dense symmetric n=2000, Lanczos m=100, lowest 5
  cfg 2  2D Laplacian 4096^2 CSR (16.8M rows), Lanczos m=100, lowest 5
  cfg 3  3D convection-diffusion 256^3 CSR (16.8M rows), Arnoldi m=50 (complex Scalar: the reference's class does
         not compile for real Scalar, arnoldi.hpp:857,864), 5 largest |lambda|
  cfg 4  Heisenberg ring L=24 as explicit CSR (16.8M rows): ground state with the reference's stop rule
         (tolerance 1e-12, index 0, maxIterations 200) and the fixed m=100 throughput run
  cfg 5  matrix-free Heisenberg ring, Lanczos m=40, at L=24 (2^30 states need 369 GB of basis: not runnable on a host)
with the same seeded start vectors bench.py uses.  Only small outputs are kept (alpha/beta or the Hessenberg matrix,
Ritz values, convergence-log tails, iteration counts, logs, 256 sampled eigenvector components), so bench.py and
the GPU tests can compare full-size runs with the reference on the GPU box, where /root/reference does not exist.
Wall-clock seconds and the thread count are recorded for information only.

usage: python tests/golden/make_ref_fullsize.py [1] [2] [3] [4] [5]     (CPU only, ~15 GB RAM, tens of minutes)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from cmpt_eigenex_b200 import synthetic as syn  # noqa: E402
from oracle import core, ref  # noqa: E402

SAMPLE = 256


def sample_rows(n):
    return (np.arange(SAMPLE, dtype=np.int64) * 2654435761) % n


def save(cfg, es, seconds, extra=None, kind="lanczos"):
    out = {"iterations": np.int64(es.iterations), "log": np.array("\n".join(es.log)), "eigenvalues": np.asarray(es.eigenvalues),
           "seconds": np.float64(seconds), "threads": np.int64(ref.num_threads())}
    if kind == "lanczos":
        out["alpha"], out["beta"] = es.alpha_beta()
    else:
        out["hessenberg"] = np.asarray(es.hessenberg)
        out["residue"] = np.float64(es.residue)
    for idx, series in es.convergence_log.items():
        out["convlog_%d" % idx] = np.asarray(series)
    X = np.asarray(es.eigenvectors)
    if X.size:
        rows = sample_rows(X.shape[0])
        out["sample_rows"] = rows
        out["eigenvector_samples"] = X[rows, :]
    out.update(extra or {})
    path = os.path.join(HERE, "ref_full_cfg%s.npz" % cfg)
    np.savez_compressed(path, **out)
    print("cfg %s: %d iterations in %.1f s on %d threads; eigenvalues %s -> %s" % (
        cfg, es.iterations, seconds, ref.num_threads(), np.asarray(es.eigenvalues)[:5], os.path.basename(path)), flush=True)


def lanczos(op, x0, m, nev, **kw):
    es = ref.LanczosEigenSolver("d")
    es.set_matrix_multiplication(op)
    es.init = x0
    if m is not None:
        es.min_iterations = es.max_iterations = m
    es.max_eigenvalues = nev
    es.indices_for_convergence = list(range(nev))
    for k, v in kw.items():
        setattr(es, k, v)
    t0 = time.perf_counter()
    es.compute()
    return es, time.perf_counter() - t0


def main():
    want = [a for a in sys.argv[1:]] or ["1", "2", "3", "4", "5"]
    ref.build()
    threads = os.cpu_count() or 1
    core.set_num_threads(threads)
    ref.set_num_threads(threads)
    if "1" in want:
        A = syn.dense_symmetric(2000, seed=1)
        es, dt = lanczos(core.Operator.dense(A), syn.start_vector(2000, seed=7), 100, 5)
        save("1", es, dt)
    if "2" in want:
        N = 4096
        rp, c, v = syn.laplacian2d_csr(N)
        es, dt = lanczos(core.Operator.csr(rp, c, v), syn.start_vector(N * N, seed=7), 100, 5)
        save("2", es, dt)
        del rp, c, v, es
    if "3" in want:
        M = 256
        rp, c, v = syn.convdiff3d_csr(M)
        es = ref.ArnoldiEigenSolver("z")
        es.set_matrix_multiplication(core.Operator.csr(rp, c, v.astype(complex)))
        es.init = syn.start_vector(M ** 3, seed=7).astype(complex)
        es.min_iterations = es.max_iterations = 50
        es.max_eigenvalues = 5
        es.indices_for_convergence = [0, 1, 2, 3, 4]
        t0 = time.perf_counter()
        es.compute()
        save("3", es, time.perf_counter() - t0, kind="arnoldi")
        del rp, c, v, es
    if "4" in want:
        L = 24
        rp, c, v = syn.heisenberg_csr(L)
        op = core.Operator.csr(rp, c, v)
        x0 = syn.start_vector(1 << L, seed=7)
        es, dt = lanczos(op, x0, None, 1, max_iterations=200, compute_eigenvectors_on=False)
        save("4_stop", es, dt)
        es, dt = lanczos(op, x0, 100, 5, compute_eigenvectors_on=False)
        save("4", es, dt)
        del rp, c, v, es, op
    if "5" in want:
        L = 24
        es, dt = lanczos(core.Operator.heisenberg(L, 1.0, True, "d"), syn.start_vector(1 << L, seed=7), 40, 1,
                         compute_eigenvectors_on=False)
        save("5_L24", es, dt)


if __name__ == "__main__":
    main()
