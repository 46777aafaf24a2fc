"""Multi-rank paths.  CPU: world_size-2 over gloo (partition, halo plans, id broadcast).  GPU: 2 ranks on 2
B200s against the oracle (skipped on a 1-GPU box)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from cmpt_eigenex_b200 import capi
from cmpt_eigenex_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, *args, timeout=600):
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "dist_worker.py"), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


def test_partition_and_halo_plan_single_process():
    L = capi.lib()
    n, P = 1000, 8
    b = [L.cmb_partition_begin(n, P, q) for q in range(P + 1)]
    assert b[0] == 0 and b[-1] == n and all(b[i] < b[i + 1] for i in range(P))
    # Heisenberg shard: remote columns are found, sorted, grouped by owner, and the remap is consistent
    Ls, P, rank = 10, 4, 2
    n = 1 << Ls
    r0, r1 = L.cmb_partition_begin(n, P, rank), L.cmb_partition_begin(n, P, rank + 1)
    rp, c, v = syn.heisenberg_csr(Ls, r0=r0, r1=r1)
    col_local = np.empty_like(c)
    cnt, per = C.c_int64(), np.zeros(P, np.int64)
    halo = np.empty(c.size, np.int32)
    capi.check(L.cmb_plan_halo(n, P, rank, c.size, capi.ptr(c), capi.ptr(col_local), C.byref(cnt), capi.ptr(per),
                               capi.ptr(halo), halo.size))
    halo = halo[: cnt.value]
    remote = np.unique(c[(c < r0) | (c >= r1)])
    assert np.array_equal(halo, remote)
    owners = np.searchsorted(np.array([L.cmb_partition_begin(n, P, q) for q in range(1, P + 1)]), halo, side="right")
    assert np.array_equal(np.bincount(owners, minlength=P), per) and per[rank] == 0
    own = (c >= r0) & (c < r1)
    assert np.array_equal(col_local[own], c[own] - r0)
    assert np.array_equal(halo[col_local[~own] - (r1 - r0)], c[~own])
    # out-of-range column is rejected
    bad = c.copy()
    bad[0] = n + 5
    assert L.cmb_plan_halo(n, P, rank, bad.size, capi.ptr(bad), None, None, None, None, 0) != 0


def test_world_size_2_gloo_cpu():
    p = _torchrun(2, "--cpu")
    assert p.returncode == 0 and "DIST_CPU_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """bench.py --impl reference launched like the driver does for N > 1: rank 0 alone runs the reference's own CPU
    implementation (oracle/_ref) on the workload and prints ONE JSON line, the other rank exits 0 without output
    (small grid here to keep the suite fast)."""
    import json

    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref/libref.so absent and no reference tree to build it from")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29633", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "2", "--warmup", "0", "--grid", "96", "--krylov-m", "30"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["metric"] == "krylov_iterations_per_sec"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["extrapolated"] is False and "2 full solve(s)" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["m"] == 30 and d["config"]["n"] == 96 * 96


def test_reference_arm_extrapolates_only_when_a_full_solve_does_not_fit():
    """With a tiny time budget the reference arm times a shorter solve, scales it with the step-cost model and says so."""
    import json

    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref/libref.so absent and no reference tree to build it from")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--grid", "128",
           "--krylov-m", "60", "--ref-budget", "1e-6"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    d = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["extrapolated"] is True and "EXTRAPOLATED" in d["cpu_baseline"]["sample"]
    assert d["value"] > 0


@pytest.mark.gpu
def test_two_ranks_on_two_gpus_match_oracle():
    n = C.c_int(0)
    capi.check(capi.lib().cmb_device_count(C.byref(n)))
    if n.value < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    p = _torchrun(2)
    assert p.returncode == 0 and "DIST_GPU_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-6000:]


def test_bench_parity_gate_accepts_rounding_and_rejects_a_wrong_answer():
    """The gate bench.py applies to every timed run (it exits 3 when `ok` is false): rounding-level differences pass,
    a perturbation of 1e-8 in one coefficient or eigenvalue fails; for Arnoldi only the converged Ritz values count."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(0)
    want = {"alpha": rng.normal(size=101) + 4, "beta": np.abs(rng.normal(size=100)) + 1, "eigenvalues": np.sort(rng.normal(size=5))}
    noise = lambda a, s: a + s * rng.normal(size=a.shape)  # noqa: E731
    good = {k: noise(v, 1e-13) for k, v in want.items()}
    p = bench.parity_against(good, want, "unit")
    assert p["ok"] and p["worst"] < 1e-11 and p["iterations_compared"] == 100
    bad = {k: v.copy() for k, v in good.items()}
    bad["beta"][37] += 1e-8
    assert not bench.parity_against(bad, want, "unit")["ok"]
    bad = {k: v.copy() for k, v in good.items()}
    bad["eigenvalues"][0] *= 1 + 1e-8
    assert not bench.parity_against(bad, want, "unit")["ok"]
    # a prefix comparison (live cpu_baseline run of the first iterations) ignores what lies beyond it
    bad = {k: v.copy() for k, v in good.items()}
    bad["alpha"][60] += 1.0
    assert bench.parity_against(bad, want, "unit", count=20)["ok"]
    # Arnoldi: the projected matrix and the converged Ritz values are gated, unconverged Ritz values are not
    H = np.triu(rng.normal(size=(12, 12)), -1)
    ev = rng.normal(size=5) + 1j * rng.normal(size=5)
    want = {"hessenberg": H, "eigenvalues": ev}
    got = {"hessenberg": H + 1e-14, "eigenvalues": ev * np.array([1, 1, 1, 1 + 1e-6, 1]), "ritz_residuals": np.array([1e-12, 1e-12, 1e-12, 1e-3, 1e-12])}
    assert bench.parity_against(got, want, "unit")["ok"]
    got["ritz_residuals"][3] = 1e-12
    assert not bench.parity_against(got, want, "unit")["ok"]
    got = {"hessenberg": H.copy(), "eigenvalues": ev, "ritz_residuals": np.full(5, 1e-12)}
    got["hessenberg"][2, 3] += 1e-7
    assert not bench.parity_against(got, want, "unit")["ok"]
