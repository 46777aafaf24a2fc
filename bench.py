#!/usr/bin/env python
"""bench.py — Krylov iterations/s of the Lanczos hot path (BASELINE.json metric) on N B200s.

Workload (config.workload): BASELINE cfg 2 — Lanczos with full reorthogonalisation on the 2D 5-point
Laplacian, 4096 x 4096 grid (16.8M rows, 83.9M non-zeros, CSR -> SELL-32 on the device), m = 100
(reference semantics: setMinIterations(m); setMaxIterations(m) => 101 vectors, 101 operator applies),
lowest 5 eigenpairs, explicit start vector (splitmix64 seed 7).  At N > 1 the same problem is
row-partitioned over the ranks (strong scaling).

One "step" = one full solve (compute()).
  value : m * steps / device time of the Krylov loops (operator already in HBM; start-vector upload
          and the host Ritz solves are inside the timed region), CUDA events on the library's stream,
          max over ranks.
  e2e   : the same metric through the solver API with HOST buffers: CSR arrays and start vector in
          pinned host memory -> operator build (H2D + SELL conversion) -> compute() with the 5 Ritz
          vectors -> eigenvalues/eigenvectors back on the host.  Wall clock around synchronised calls.
  roofline : the dominant kernel family (CGS2 passes), algorithmic bytes per launch / mean launch
          duration from CUDA events recorded around every launch inside the timed region.
  cpu_baseline / --impl reference : the CPU oracle (restatement of the reference's algorithm:
          single-pass MGS as separate dot/axpy sweeps, tridiagonal solve every trip; the reference
          itself cannot be built here, Eigen3 is absent) on the box's host cores, on a bounded sample
          of the same workload (same matrix and start vector, the first m_sample iterations).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "krylov_iterations_per_sec"
UNIT = "it/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=4096, help="N of the N x N Laplacian grid")
    ap.add_argument("--m", type=int, default=100, help="Lanczos iterations per solve")
    ap.add_argument("--nev", type=int, default=5)
    ap.add_argument("--cpu-sample-m", type=int, default=20, help="iterations of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_name(args):
    return "cfg2: Lanczos full-reorth, 2D 5-pt Laplacian %dx%d CSR (n=%d), m=%d, lowest %d" % (
        args.grid, args.grid, args.grid * args.grid, args.m, args.nev)


# -------------------------------------------------------------------------------------------------
# CPU oracle legs
# -------------------------------------------------------------------------------------------------
def oracle_sample(args, rp, c, v, x0, m_sample, steps, warmup):
    """Times the CPU oracle on the first m_sample Lanczos iterations of the workload."""
    from oracle import core
    from oracle import reference_solvers as rs

    threads = os.cpu_count() or 1
    core.set_num_threads(threads)
    opr = core.Operator.csr(rp, c, v)
    times = []
    for i in range(warmup + steps):
        ref = rs.LanczosEigenSolver("d")
        ref.set_matrix_multiplication(opr)
        ref.init = x0
        ref.min_iterations = ref.max_iterations = m_sample
        ref.max_eigenvalues = args.nev
        ref.indices_for_convergence = list(range(args.nev))
        ref.compute_eigenvectors_on = False
        t0 = time.perf_counter()
        ref.compute()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": m_sample * len(times) / total, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "same matrix/start vector, first %d of %d Lanczos iterations (single-pass MGS as the reference, "
                      "eigenvectors off), %d timed solve(s), OpenMP over all host threads" % (m_sample, args.m, len(times)),
            "ms_per_step": 1e3 * total / len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cmpt_eigenex_b200 import synthetic as syn

    N = args.grid
    rp, c, v = syn.laplacian2d_csr(N)
    x0 = syn.start_vector(N * N, seed=7)
    # bound the whole run to a few minutes: shrink the per-step sample when many steps are requested
    m_sample = args.cpu_sample_m
    nrun = args.steps + args.warmup
    while m_sample > 4 and nrun * (m_sample ** 2) > 6 * 20 ** 2:
        m_sample -= 2
    cb = oracle_sample(args, rp, c, v, x0, m_sample, args.steps, args.warmup)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "note": "CPU oracle port of the reference algorithm; the "
                       "reference itself needs Eigen3, which is not installed"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import cmpt_eigenex_b200 as pkg
    from cmpt_eigenex_b200 import capi
    from cmpt_eigenex_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from cmpt_eigenex_b200 import dist

        ctx = dist.make_context(local_rank)
    else:
        ctx = pkg.Context(0)
        dist = None

    N, m, nev = args.grid, args.m, args.nev
    n = N * N
    r0, r1 = (rank * n) // world, ((rank + 1) * n) // world
    nloc = r1 - r0
    rp, c, v = syn.laplacian2d_csr(N, r0, r1)
    nnz_local = int(rp[-1])
    # inputs staged in pinned host memory (what the e2e region copies from)
    prp, pc, pv = capi.PinnedBuffer(rp.shape, np.int64), capi.PinnedBuffer(c.shape, np.int32), capi.PinnedBuffer(v.shape, np.float64)
    prp.array[:], pc.array[:], pv.array[:] = rp, c, v
    x0_full = syn.start_vector(n, seed=7)
    px = capi.PinnedBuffer((nloc,), np.float64)
    px.array[:] = x0_full[r0:r1]
    del rp, c, v

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        return dist.all_max(x) if dist is not None else x

    def make_solver(op, vectors):
        es = pkg.LanczosEigenSolver(np.float64)
        if op is not None:
            es.setMatrixMultiplication(op)
        es.setInitialVector(px.array)
        es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(nev)
        es.setIndicesForConvergence(list(range(nev))).setComputeEigenvectorsOn(vectors).setReserveSize(m + 1)
        return es

    # ---- device-resident leg: operator in HBM, Krylov loop timed with CUDA events ----
    op = pkg.DeviceOperator.from_csr(ctx, prp.array, pc.array, pv.array, n_global=n, row_begin=r0)
    es = make_solver(op, vectors=False)
    for _ in range(args.warmup):
        es.compute()
    ctx.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ctx.profile(True)
    dev_ms = 0.0
    host_ms = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()  # L2 flushed between timed iterations (inputs are also far larger than L2)
        ctx.sync()
        barrier()
        t_host = time.perf_counter()
        ctx.timer_start()
        es.compute()
        dev_ms += ctx.timer_stop()
        host_ms += (time.perf_counter() - t_host) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0  # kernels of this library launched inside the timed region
    fams = {}
    for fam in ("cgs_dot", "cgs_update_dot", "cgs_update_norm", "spmv_sell", "vec_dot", "nccl_allreduce", "nccl_halo",
                "halo_pack"):
        fams[fam] = ctx.profile_get(fam)
    ctx.profile(False)
    dev_ms = max_over_ranks(dev_ms)
    if os.environ.get("BENCH_DEBUG") and rank == 0:
        print("device-leg: dev %.1f ms, host wall %.1f ms per solve; families %s" % (
            dev_ms / args.steps, host_ms / args.steps, {k: (round(v[0] / args.steps, 2), v[1] // args.steps) for k, v in fams.items()}),
            file=sys.stderr, flush=True)
    value = m * args.steps / (dev_ms * 1e-3)
    eig = es.eigenvalues()
    step_bytes = es.deviceBytes()  # algorithmic bytes of one solve on this rank
    es.close()

    # ---- roofline of the dominant kernel family ----
    peak, peak_src = load_peaks()
    s = 8.0
    alg = {"cgs_dot": sum((cc + 1) for cc in range(1, m + 1)) * nloc * s,
           "cgs_update_dot": sum((cc + 2) for cc in range(1, m + 1)) * nloc * s,
           "cgs_update_norm": sum((cc + 2) for cc in range(1, m + 1)) * nloc * s,
           "spmv_sell": (m + 1) * op.bytes}
    dom = max(alg, key=lambda k: fams[k][0])
    dom_ms, dom_n = fams[dom]
    roof = None
    if dom_n:
        per_launch_bytes = alg[dom] * args.steps / dom_n
        per_launch_s = dom_ms * 1e-3 / dom_n
        ach = per_launch_bytes / per_launch_s / 1e9
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[dom]
            traffic = tr["ratio"] * per_launch_bytes  # measured dram/algorithmic ratio x this run's bytes per launch
            traffic_src = "ncu --set full capture (profiles/r1d_prof_step4_raw.txt): dram bytes / algorithmic bytes = %.4f" % tr["ratio"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "peak_note": "the measured peak is a copy (half reads, half writes); the CGS passes read c+1 vectors per "
                             "vector written, so frac can exceed 1 — ncu's DRAM peak on this part is about 8.19 TB/s",
                "bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_s * 1e3,
                "families": {k: {"ms": fams[k][0], "launches": fams[k][1],
                                 "GBps": (alg[k] * args.steps / (fams[k][0] * 1e-3) / 1e9) if (k in alg and fams[k][0] > 0) else None}
                             for k in fams}}
    overall_gbs = step_bytes * args.steps / (dev_ms * 1e-3) / 1e9

    # ---- end-to-end leg: host buffers in, host results out ----
    e2e = None
    if not args.no_e2e:
        op.close()
        es2 = make_solver(None, vectors=True)
        h2d = prp.array.nbytes + pc.array.nbytes + pv.array.nbytes + px.array.nbytes
        d2h = nev * nloc * 8 + nev * 8
        times = []
        for i in range(args.warmup + args.steps):
            ctx.sync()
            barrier()
            t0 = time.perf_counter()
            op2 = pkg.DeviceOperator.from_csr(ctx, prp.array, pc.array, pv.array, n_global=n, row_begin=r0)
            es2.setMatrixMultiplication(op2).setInitialVector(px.array)
            es2.compute()
            ev = es2.eigenvalues()
            X = es2.eigenvectors(copy=False)
            chk = float(X[0, 0]) + float(ev[0])  # touch the host results
            op2.close()
            ctx.sync()
            dt = time.perf_counter() - t0
            barrier()
            if i >= args.warmup:
                times.append(dt)
            if os.environ.get("BENCH_DEBUG"):
                print("e2e iter %d: %.1f ms" % (i, dt * 1e3), file=sys.stderr, flush=True)
            del chk, X
        tot = max_over_ranks(sum(times))
        e2e = {"value": m * len(times) / tot, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * tot / len(times),
               "what": "pinned host CSR + start vector -> SELL build -> compute() with %d Ritz vectors -> host" % nev}
        es2.close()

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = oracle_sample(args, prp.array, pc.array, pv.array, x0_full, args.cpu_sample_m, 1, 0)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args), "partition": "rows/%d" % world,
                           "timing": "CUDA events on the library stream; L2 flushed (256 MiB memset) between timed solves; "
                                     "basis (13.6 GB) and matrix (1.3 GB) far exceed the 126 MB L2",
                           "algorithmic_GBps": overall_gbs, "bytes_per_solve": step_bytes,
                           "lowest_eigenvalues": [float(x) for x in eig]},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
