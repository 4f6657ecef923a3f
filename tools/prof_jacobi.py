"""Minimal driver for `ncu`: a few fused Jacobi sweeps (or SpMVs) on 3D Poisson n^3 level 0 with the kernel family the
environment selects (SPARSH_PATTERN=1 -> csr-pattern8, SPARSH_DICT=0 -> plain CSR, default csr-dict16).
    python tools/prof_jacobi.py [--n 256] [--sweeps 6] [--kind K --tl T]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sparsh_amg_b200 as sp  # noqa: E402
from sparsh_amg_b200 import generators  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--sweeps", type=int, default=6)
    ap.add_argument("--kind", type=int, default=-1)
    ap.add_argument("--tl", type=int, default=128)
    a = ap.parse_args()
    sp.init(0)
    A = generators.poisson_7pt(a.n, a.n, a.n)
    dA = sp.DeviceMatrix.from_csr(A)
    if a.kind >= 0:
        dA.force_kernel(a.kind, a.tl)
    print("kernel:", dA.kernel(), flush=True)
    rng = np.random.default_rng(0)
    x, b, t = (sp.DeviceVector(data=rng.standard_normal(A.nrow)) for _ in range(3))
    lib = sp.capi.load()
    sp.capi.check(lib.sparsh_jacobi(dA.h, b.ptr, x.ptr, t.ptr, 0.66667, a.sweeps))
    sp.capi.check(lib.sparsh_spmv(dA.h, x.ptr, t.ptr))
    sp.capi.check(lib.sparsh_sync())
    print("done", flush=True)


if __name__ == "__main__":
    main()
