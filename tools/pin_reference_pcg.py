"""Pin the full-size AMG-PCG solve to the reference itself (developer tool; needs oracle/_ref, i.e. /root/reference).

Runs the reference's STOCK Solver_PCG_1 (src/AMG_main_solvers.cpp:107-167, unmodified sources compiled by oracle/Makefile
against the OpenMP MKL shim) on 3D 7-point Poisson n^3, b = 1, x0 = 0, to ||r|| <= 1e-8 * ||b|| (the reference's absolute
tolerance `tol1` is a run-time hook in the generated header, set to that value), and freezes the residual history its own
prints report into tests/golden/pcg_poisson3d_<n>_ref.json.  The -m gpu tests compare the CUDA path's history at the same
size with it (1e-10 relative, count +-1); bench.py's reference arm quotes its iteration count.

    python tools/pin_reference_pcg.py 256 [threads]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_bindings import Oracle, Ref  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    o, r = Oracle.get(), Ref.get()
    o.set_threads(threads)
    r.set_threads(threads)
    A = o.gen_poisson3d(n, n, n)
    b = np.ones(A.nrow)
    tol = 1e-8 * float(np.linalg.norm(b))
    r.set_tol(tol)
    t0 = time.time()
    x, hist = r.solve("Solver_PCG_1", A, b, np.zeros(A.nrow))
    secs = time.time() - t0
    res = float(np.linalg.norm(b - A.to_scipy() @ x))
    out = {"grid": n, "rows": int(A.nrow), "nnz": int(A.nnz), "tol_abs": tol, "iterations": len(hist),
           "initial_residual": float(np.linalg.norm(b)),  # x0 = 0; the reference prints only ||r|| AFTER each iteration
           "history_after_iteration": [float(v) for v in hist], "true_residual": res, "threads": threads,
           "setup_plus_solve_seconds": secs,
           "how": "reference Solver_PCG_1 (stock, unmodified sources + OpenMP MKL shim), history parsed from its own prints"}
    path = os.path.join(ROOT, "tests", "golden", f"pcg_poisson3d_{n}_ref.json")
    json.dump(out, open(path, "w"), indent=1)
    nb = float(np.linalg.norm(b))
    print(f"{n}^3: {len(hist)} iterations, ||r||/||b|| = {hist[-1]/nb:.3e}, true {res/nb:.3e}, {secs:.1f}s -> {path}")


if __name__ == "__main__":
    main()
