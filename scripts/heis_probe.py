"""Times the matrix-free Heisenberg apply (cfg 5 operator) per launch family, with and without the sibling bonds.

usage: python scripts/heis_probe.py [L] [mode ...]   (0 = passes stay inside their tiles, 1 = sibling bonds)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cmpt_eigenex_b200 as pkg  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 27
settings = [int(a) for a in sys.argv[2:]] or [0, 1]
ctx = pkg.Context(0)
x = np.random.default_rng(1).standard_normal(1 << L)
x /= np.linalg.norm(x)
for ncl in settings:
    os.environ["CMPT_B200_HEIS_SIBLINGS"] = "1" if ncl else "0"
    op = pkg.DeviceOperator.heisenberg(ctx, L)
    op.apply(x)  # plans the passes, warms up
    ctx.profile(True)
    for _ in range(3):
        op.apply(x)
    ms, n = ctx.profile_get("heisenberg_mf")
    ctx.profile(False)
    print("L=%d siblings=%s: %d launches, %.3f ms per apply, %.3f ms per launch" % (L, "on" if ncl else "off", n, ms / 3, ms / max(n, 1)), flush=True)
    op.close()
