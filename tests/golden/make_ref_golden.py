"""Generates tests/golden/ref_traces.npz from the reference ITSELF.

Every case of tests/ref_cases.py is run through oracle/_ref/libref.so — the unmodified
/root/reference/include/cmpt/eigen_ex/{lanczos,arnoldi}.hpp compiled against oracle/eigen_shim/ (oracle/Makefile,
target `ref`) — and alpha/beta (or the Hessenberg matrix), eigenvalues, eigenvectors, iteration count, log strings
and convergence logs are recorded.  The file travels to the GPU box, where /root/reference does not exist.

usage: python tests/golden/make_ref_golden.py      (CPU only; needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_cases  # noqa: E402
from oracle import core, ref  # noqa: E402


def main():
    if not ref.reference_present():
        raise SystemExit("needs the reference tree at " + ref.REFERENCE)
    ref.build(force=True)
    core.set_num_threads(1)
    ref.set_num_threads(1)
    out = {}
    for case in ref_cases.CASES:
        es = ref_cases.run_checker(ref, case)
        for k, v in ref_cases.record(es, case).items():
            out["%s/%s" % (case, k)] = v
        print("%-28s iterations %4d  eigenvalues %s" % (case, es.iterations, np.asarray(es.eigenvalues)[:3]))
    np.savez_compressed(os.path.join(HERE, "ref_traces.npz"), **out)
    print("wrote ref_traces.npz: %d arrays, %.1f KB" % (len(out), os.path.getsize(os.path.join(HERE, "ref_traces.npz")) / 1e3))


if __name__ == "__main__":
    main()
