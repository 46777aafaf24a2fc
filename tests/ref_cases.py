"""Case table shared by the reference-pinned fixtures (tests/golden/ref_traces.npz), the CPU pin tests and the
GPU parity tests.  Each case is a call sequence on a seeded operator, run three ways:

  * through the UNMODIFIED reference classes (oracle/_ref, oracle/ref.py)      -> recorded in ref_traces.npz
  * through the restatement (oracle/reference_solvers.py)                     -> CPU test against the fixture
  * through the product (cmpt_eigenex_b200, CUDA via the C-ABI)               -> GPU test against the fixture and,
                                                                                 when libref.so is present, live

The first three cases are the reference's own samples (src/samples/sample_lanczos1.cpp, sample_lanczos2.cpp,
sample_arnoldi.cpp) with their exact settings; the rest are small instances of BASELINE.json's configs.
"""
import numpy as np

from cmpt_eigenex_b200 import synthetic as syn


def _dense_sample1():
    return np.array([[1.0, 0.5, 0.0], [0.5, 2.0, 0.5], [0.0, 0.5, 3.0]])


def _random_complex(n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))


# name -> dict(kind, prefix, op=(type, args...), init, settings, [second leg for continueToCompute])
CASES = {
    # sample_lanczos1.cpp:13-32
    "sample_lanczos1": dict(kind="lanczos", p="d", op=("dense", _dense_sample1), init=None,
                            settings=dict(tolerance=1.0e-5, max_iterations=100)),
    # sample_lanczos2.cpp:22-59
    "sample_lanczos2": dict(kind="lanczos", p="z", op=("csr", lambda: syn.hermitian_chain_csr(200)), init=("seeded", 1),
                            settings=dict(shift=0.0, tolerance=1.0e-7, threshold=1.0e-14, min_iterations=-1, max_iterations=1000,
                                          max_eigenvalues=10, indices_for_convergence=[0], interval=1),
                            stop_rule=True),
    # sample_arnoldi.cpp:21-41 (Eigen's MatrixType::Random is rand()-based and not reproducible across libraries:
    # the matrix is a seeded uniform(-1,1) complex one of the same shape)
    "sample_arnoldi": dict(kind="arnoldi", p="z", op=("dense", lambda: _random_complex(50)), init=None,
                           settings=dict(min_iterations=40, max_iterations=40, tolerance=1.0e-14, max_eigenvalues=2)),
    # cfg 1 shape: dense symmetric, lowest 5
    "dense300_m60": dict(kind="lanczos", p="d", op=("dense", lambda: syn.dense_symmetric(300, seed=1)), init=("splitmix", 7),
                         settings=dict(min_iterations=60, max_iterations=60, max_eigenvalues=5, indices_for_convergence=[0, 1, 2, 3, 4])),
    # cfg 2 shape: 2D Laplacian CSR, m = 100, full reorthogonalisation
    "laplacian48_m100": dict(kind="lanczos", p="d", op=("csr", lambda: syn.laplacian2d_csr(48)), init=("splitmix", 7),
                             settings=dict(min_iterations=100, max_iterations=100, max_eigenvalues=5,
                                           indices_for_convergence=[0, 1, 2, 3, 4])),
    # cfg 3 shape: convection-diffusion CSR, Arnoldi m = 50, complex Scalar (the reference's class needs it)
    "convdiff8_m30": dict(kind="arnoldi", p="z", op=("csr", lambda: syn.convdiff3d_csr(8)), init=("splitmix", 7),
                          settings=dict(min_iterations=30, max_iterations=30, max_eigenvalues=5, indices_for_convergence=[0, 1, 2])),
    # cfg 4 semantics: Heisenberg ring as explicit CSR, ground state with the reference's stop rule
    "heisenberg12_stop": dict(kind="lanczos", p="d", op=("csr", lambda: syn.heisenberg_csr(12)), init=("splitmix", 7),
                              settings=dict(max_iterations=200, max_eigenvalues=1), stop_rule=True),
    # cfg 5 shape: matrix-free Heisenberg chain, fixed m = 40
    "heisenberg_mf12_m40": dict(kind="lanczos", p="d", op=("heisenberg", 12, 1.0, True), init=("splitmix", 7),
                                settings=dict(min_iterations=40, max_iterations=40, max_eigenvalues=2, indices_for_convergence=[0, 1])),
    # shift + continueToCompute (lanczos.hpp:701-712)
    "laplacian20_shift_continue": dict(kind="lanczos", p="d", op=("csr", lambda: syn.laplacian2d_csr(20)), init=("splitmix", 11),
                                       settings=dict(shift=0.3, min_iterations=30, max_iterations=30, max_eigenvalues=3,
                                                     indices_for_convergence=[0, 1]),
                                       then=dict(min_iterations=55, max_iterations=55)),
    # complex Hermitian with a strided reorthogonalisation interval is rounding-sensitive; keep interval 1 but complex + deflation
    "chain60_deflated": dict(kind="lanczos", p="z", op=("csr", lambda: syn.hermitian_chain_csr(60)), init=("seeded", 3),
                             settings=dict(min_iterations=25, max_iterations=25, max_eigenvalues=3), deflate=1),
}


def operator_arrays(case):
    op = CASES[case]["op"]
    if op[0] == "heisenberg":
        return op
    return (op[0], op[1]())


def start_vector(case, n):
    init = CASES[case]["init"]
    if init is None:
        return None
    p = CASES[case]["p"]
    if init[0] == "splitmix":
        x = syn.start_vector(n, seed=init[1])
        if p == "z":
            x = x + 1j * syn.start_vector(n, seed=init[1] + 100)
        return x
    if init[0] == "seeded":  # makeRandomVector(std::mt19937(seed), n): libstdc++ stream, taken from the oracle library
        from oracle import core

        return core.seeded_vector(init[1], n, p)
    raise ValueError(init)


def deflation_vectors(case, n):
    k = CASES[case].get("deflate", 0)
    p = CASES[case]["p"]
    rng = np.random.default_rng(17)
    out = []
    for _ in range(k):
        w = rng.standard_normal(n) + (1j * rng.standard_normal(n) if p == "z" else 0)
        for d in out:
            w = w - np.vdot(d, w) * d
        out.append(w / np.linalg.norm(w))
    return out


# ---------------------------------------------------------------------------------------------------------
# running a case through a checker module (oracle.reference_solvers or oracle.ref: same Python surface)
# ---------------------------------------------------------------------------------------------------------
def checker_operator(case):
    from oracle import core

    p = CASES[case]["p"]
    arr = operator_arrays(case)
    if arr[0] == "dense":
        return core.Operator.dense(arr[1].astype(complex) if p == "z" else arr[1])
    if arr[0] == "csr":
        rp, c, v = arr[1]
        return core.Operator.csr(rp, c, v.astype(complex) if p == "z" else v)
    return core.Operator.heisenberg(arr[1], arr[2], arr[3], p)


def run_checker(mod, case):
    cs = CASES[case]
    op = checker_operator(case)
    es = (mod.LanczosEigenSolver if cs["kind"] == "lanczos" else mod.ArnoldiEigenSolver)(cs["p"])
    es.set_matrix_multiplication(op)
    x0 = start_vector(case, op.n)
    if x0 is not None:
        es.init = x0
    es.ortho = deflation_vectors(case, op.n)
    for k, v in cs["settings"].items():
        setattr(es, k, v)
    es.compute()
    if "then" in cs:
        for k, v in cs["then"].items():
            setattr(es, k, v)
        es.continue_to_compute()
    return es


def record(es, case):
    """Everything a fixture keeps about a checker run."""
    cs = CASES[case]
    out = {"iterations": np.int64(es.iterations), "log": np.array("\n".join(es.log)), "eigenvalues": np.asarray(es.eigenvalues),
           "eigenvectors": np.asarray(es.eigenvectors), "init": np.asarray(es.init)}
    if cs["kind"] == "lanczos":
        out["alpha"], out["beta"] = es.alpha_beta()
    else:
        out["hessenberg"] = np.asarray(es.hessenberg)
    for idx in es.indices_for_convergence:
        if idx in es.convergence_log:
            out["convlog_%d" % idx] = np.asarray(es.convergence_log[idx])
    return out


# ---------------------------------------------------------------------------------------------------------
# running a case through the product (GPU)
# ---------------------------------------------------------------------------------------------------------
def run_product(pkg, ctx, case, init=None):
    cs = CASES[case]
    dt = np.complex128 if cs["p"] == "z" else np.float64
    arr = operator_arrays(case)
    if arr[0] == "dense":
        op = pkg.DeviceOperator.from_dense(ctx, arr[1].astype(dt))
    elif arr[0] == "csr":
        rp, c, v = arr[1]
        op = pkg.DeviceOperator.from_csr(ctx, rp, c, v.astype(dt))
    else:
        op = pkg.DeviceOperator.heisenberg(ctx, arr[1], arr[2], arr[3], dtype=dt)
    es = (pkg.LanczosEigenSolver if cs["kind"] == "lanczos" else pkg.ArnoldiEigenSolver)(dt)
    es.setMatrixMultiplication(op)
    n = op.height
    x0 = init if init is not None else start_vector(case, n)
    if x0 is not None:
        es.setInitialVector(x0)
    defl = deflation_vectors(case, n)
    if defl:
        es.setOrthogonalizingVectors(defl)
    setters = {"tolerance": "setTolerance", "threshold": "setThreshold", "shift": "setEigenvalueShift",
               "min_iterations": "setMinIterations", "max_iterations": "setMaxIterations", "max_eigenvalues": "setMaxEigenvalues",
               "indices_for_convergence": "setIndicesForConvergence", "interval": "setReorthogonalizeInterval"}

    def apply(settings):
        for k, v in settings.items():
            getattr(es, setters[k])(v)

    apply(cs["settings"])
    es.compute()
    if "then" in cs:
        apply(cs["then"])
        es.continueToCompute()
    return es, op
