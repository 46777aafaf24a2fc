"""The row-partitioned code paths on ONE GPU: P virtual ranks (cmpt_b200_debug.h) driven by P host threads.

A virtual rank is an ordinary context whose compute stream owns a disjoint share of the SMs (CUDA green context) and
whose peer buffers are plain device pointers; everything on the per-step path — gram_schmidt2_mailed (peer-memory
mailboxes), halo_push_part / halo_wait_cta fused into the SpMV, the all-push mode, the Pythagorean beta and its guard,
the slab exchange of the matrix-free Heisenberg apply, device-side breakdown — is the code real ranks run.  The checks
are the ones tests/dist_worker.py runs on real GPUs (tests/multirank_checks.py), against the restatement and, where the
library is available, against the reference itself (oracle/_ref).
"""
import os

import numpy as np
import pytest

import cmpt_eigenex_b200 as pkg
import multirank_checks as mc
from cmpt_eigenex_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

# Every section of the shared checks.  The matrix-free Heisenberg operator exchanges its slabs between virtual ranks with
# SM stores (the fused push of the Gram-Schmidt pass that produces w, and a stand-alone push kernel for the first apply);
# the copy-engine exchange real ranks use for applies without a producing pass needs SMs on one device and is not taken.
SECTIONS = ["laplacian", "heisenberg", "convdiff_arnoldi", "exhaust", "breakdown_at_k", "deflation", "one_directional",
            "heisenberg_mf"]


def _expected(which):
    from oracle import core

    core.set_num_threads(2)
    if which == "restatement":
        from oracle import reference_solvers as rs

        return mc.expected(rs, core)
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref/libref.so absent and no reference tree to build it from")
    ref.set_num_threads(2)
    return mc.expected(ref, core)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_virtual_ranks_match_restatement(nranks):
    exp = _expected("restatement")
    results, info = pkg.run_virtual_ranks(nranks, lambda ctx, comm: mc.run_checks(pkg, ctx, comm, exp, only=SECTIONS))
    assert info["nranks"] == nranks
    for r in results[1:]:
        assert r == results[0]  # every rank reports the same eigenvalues, bit for bit


def test_virtual_ranks_match_the_reference_itself():
    exp = _expected("reference")
    results, _ = pkg.run_virtual_ranks(4, lambda ctx, comm: mc.run_checks(pkg, ctx, comm, exp, only=SECTIONS))
    assert all(r == results[0] for r in results)


def test_virtual_ranks_use_green_contexts():
    g = pkg.VirtualGroup(0, 4)
    info = g.info()
    g.close()
    assert info["nranks"] == 4 and info["sms_per_rank"] >= 8
    if not info["green_contexts"]:
        pytest.skip("the driver offers no green contexts: virtual ranks share the SMs (grids sized for 1/P of the device)")
    assert info["sms_per_rank"] % 8 == 0


def test_pythagorean_beta_guard_falls_back_to_the_explicit_norm(monkeypatch):
    """CMPT_B200_NORM_GUARD raises the guard ratio so that it fires on EVERY step: each step is then halted on the
    device, re-run with the explicitly reduced norm and the chain resumed — alpha/beta must still be the checker's."""
    from oracle import core
    from oracle import reference_solvers as rs

    core.set_num_threads(2)
    N, m = 24, 30
    n = N * N
    full = syn.laplacian2d_csr(N)
    x0 = syn.start_vector(n, seed=7)
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(*full))
    ref.init = x0
    ref.min_iterations = ref.max_iterations = m
    ref.max_eigenvalues = 2
    ref.compute()
    ra, rb = ref.alpha_beta()

    def work(ctx, comm):
        r0, r1 = comm.row_range(n)
        op = pkg.DeviceOperator.from_csr(ctx, *mc.shard_of(full, r0, r1), n_global=n, row_begin=r0)
        es = pkg.LanczosEigenSolver()
        es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(2)
        es.compute()
        out = (es.alpha(), es.beta(), es.eigenvalues(), es.log())
        es.close()
        op.close()
        return out

    plain, _ = pkg.run_virtual_ranks(2, work)
    monkeypatch.setenv("CMPT_B200_NORM_GUARD", "2.0")
    guarded, _ = pkg.run_virtual_ranks(2, work)
    for a, b, ev, log in plain + guarded:
        assert np.abs(a - ra).max() < 1e-11 and np.abs(b - rb).max() < 1e-11
        assert np.abs(ev - ref.eigenvalues).max() < 1e-10 * np.abs(ra).max()
        assert log == ref.log
    # the two norms differ in the last bits only (Pythagorean identity vs explicit reduction)
    assert np.abs(plain[0][1] - guarded[0][1]).max() < 1e-13


def test_peer_wait_timeout_is_reported_and_poisons_the_context():
    """One rank never launches its share of a distributed apply: the waiting rank's kernel gives up after the (shortened)
    spin bound, the call returns CMB_ERR_NCCL instead of stale data, and the context refuses further work."""
    from cmpt_eigenex_b200 import capi

    n = 4096
    full = syn.laplacian2d_csr(64)
    x0 = syn.start_vector(n, seed=3)

    def work(ctx, comm):
        ctx.set_spin_timeout(0.5)
        r0, r1 = comm.row_range(n)
        op = pkg.DeviceOperator.from_csr(ctx, *mc.shard_of(full, r0, r1), n_global=n, row_begin=r0)
        op.apply(x0[r0:r1])  # a healthy exchange first
        comm.barrier()
        outcome = "skipped"
        if comm.rank == 0:
            try:
                op.apply(x0[r0:r1])
                outcome = "returned"
            except capi.CmbError as e:
                outcome = "error %d" % e.code
                assert "timed out" in str(e)
            with pytest.raises(capi.CmbError):
                op.apply(x0[r0:r1])  # the context is dead now
        comm.barrier()
        return outcome

    results, _ = pkg.run_virtual_ranks(2, work)
    assert results[0] == "error -4" and results[1] == "skipped"


def test_a_bad_shard_on_one_rank_fails_the_operator_on_every_rank():
    """Rank 1 hands in a column index outside the matrix: its build fails validation, and rank 0 — whose shard is fine —
    gets an error too instead of an operator whose peer does not exist (it would hang in the first exchange)."""
    from cmpt_eigenex_b200 import capi

    n = 4096
    full = syn.laplacian2d_csr(64)

    def work(ctx, comm):
        r0, r1 = comm.row_range(n)
        rp, c, v = mc.shard_of(full, r0, r1)
        if comm.rank == 1:
            c = c.copy()
            c[5] = n + 17
        try:
            pkg.DeviceOperator.from_csr(ctx, rp, c, v, n_global=n, row_begin=r0)
            return "built"
        except capi.CmbError as e:
            return "error %d: %s" % (e.code, e)

    results, _ = pkg.run_virtual_ranks(2, work)
    assert results[0].startswith("error -1") and "other rank" in results[0]
    assert results[1].startswith("error -1") and "column index" in results[1]
