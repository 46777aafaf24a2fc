"""Worker of the multi-rank tests: launched by torch.distributed.run, one process per GPU (or, with
--cpu, one process per rank on the CPU for the host-side logic only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cmpt_eigenex_b200 as pkg  # noqa: E402
from cmpt_eigenex_b200 import capi, dist  # noqa: E402
from cmpt_eigenex_b200 import synthetic as syn  # noqa: E402


def gather_rows(local, n, td):
    import torch

    world = td.get_world_size()
    parts = [None] * world
    td.all_gather_object(parts, np.asarray(local))
    return np.concatenate(parts)


def cpu_main():
    """host-side logic at world_size 2 over gloo: partition, halo plans, id broadcast."""
    import ctypes as C

    import torch

    td = dist.init()
    rank, world = td.get_rank(), td.get_world_size()
    L = capi.lib()
    N = 12
    n = N * N
    r0, r1 = dist.row_range(n)
    assert (r0, r1) == (L.cmb_partition_begin(n, world, rank), L.cmb_partition_begin(n, world, rank + 1))
    rp, c, v = syn.laplacian2d_csr(N, r0, r1)
    col_local = np.empty_like(c)
    cnt = C.c_int64()
    per = np.zeros(world, np.int64)
    halo = np.empty(c.size, np.int32)
    capi.check(L.cmb_plan_halo(n, world, rank, c.size, capi.ptr(c), capi.ptr(col_local), C.byref(cnt), capi.ptr(per),
                               capi.ptr(halo), halo.size))
    halo = halo[: cnt.value]
    # what I need is exactly the set of remote columns, owned by the other rank
    remote = np.unique(c[(c < r0) | (c >= r1)])
    assert np.array_equal(halo, remote) and per[rank] == 0 and per.sum() == halo.size
    assert halo.size == N  # one grid line of the neighbour block
    own = (c >= r0) & (c < r1)
    assert np.array_equal(col_local[own], c[own] - r0)
    assert np.array_equal(halo[col_local[~own] - (r1 - r0)], c[~own])
    # the lists are consistent across ranks: every index I ask for is owned by the peer
    lists = [None] * world
    td.all_gather_object(lists, (r0, r1, halo.tolist()))
    for q, (q0, q1, need) in enumerate(lists):
        for g in need:
            owner = [k for k, (a, b, _) in enumerate(lists) if a <= g < b]
            assert owner and owner[0] != q
    # 128-byte token broadcast (the path the NCCL id takes)
    tok = np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, np.uint8)
    t = torch.from_numpy(tok)
    td.broadcast(t, src=0)
    assert np.array_equal(t.numpy(), np.arange(128, dtype=np.uint8))
    assert dist.all_max(float(rank)) == world - 1 and dist.all_sum(1.0) == world
    td.barrier()
    if rank == 0:
        print("DIST_CPU_OK")


def gpu_main():
    """Real ranks: the shared multi-rank checks (tests/multirank_checks.py) with one process per GPU."""
    import multirank_checks as mc
    from oracle import core
    from oracle import reference_solvers as rs

    td = dist.init()
    rank, world = td.get_rank(), td.get_world_size()
    ctx = dist.make_context()
    core.set_num_threads(2)

    class Comm:
        def __init__(self):
            self.rank, self.world = rank, world

        def barrier(self):
            td.barrier()

        def gather(self, obj):
            parts = [None] * world
            td.all_gather_object(parts, obj)
            return parts

        def row_range(self, n):
            return dist.row_range(n)

    exp = mc.expected(rs, core)
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--only=")]
    results = mc.run_checks(pkg, ctx, Comm(), exp, only=only[0] if only else None)
    td.barrier()
    ctx.close()
    if rank == 0:
        print("DIST_GPU_OK", results)


if __name__ == "__main__":
    if "--cpu" in sys.argv:
        cpu_main()
    else:
        gpu_main()
