"""Runs tests/multirank_checks.run_checks on virtual ranks with a short spin bound and prints the failing location."""
import os
import sys
import traceback

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cmpt_eigenex_b200 as pkg  # noqa: E402
import multirank_checks as mc  # noqa: E402
from oracle import core  # noqa: E402
from oracle import reference_solvers as rs  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ONLY = sys.argv[2].split(",") if len(sys.argv) > 2 else None
core.set_num_threads(2)
exp = mc.expected(rs, core)


def work(ctx, comm):
    ctx.set_spin_timeout(4.0)
    try:
        return mc.run_checks(pkg, ctx, comm, exp, only=ONLY)
    except Exception:
        print("rank", comm.rank, "FAILED:\n" + traceback.format_exc(limit=4), flush=True)
        raise


try:
    print(pkg.run_virtual_ranks(P, work, timeout=200))
except Exception as e:
    print("ERROR", str(e)[:300])
