"""Measurement helpers shared by bench.py (N = 1) and distributed.bench_main (N > 1): per-kernel timing with CUDA events
on the library's stream, the byte models of SURVEY §8d, and the roofline objects of the bench line.

Byte accounting.  Every figure is per launch, n rows, z stored entries:
  * "algorithmic CSR bytes": what the operation moves on plain CSR: 12 z + 4 (n + 1) + vectors;
  * "stored bytes": what the kernel that actually runs has to move.  plain CSR: the same.  csr-dict16: 2 z + 4 (n + 1) +
    vectors (one 16-bit code per entry; the diagonal array is still read by the Jacobi epilogue).  csr-pattern8: n + vectors
    (one byte per row; row pointer, values, columns AND the Jacobi diagonal come from the shared-memory table).
`roofline.achieved` is stored bytes / time (a DRAM-roofline figure: never meaningfully above the measured copy peak);
the gain of the lossless re-encoding is reported separately as `speedup_vs_csr_bound` = algorithmic CSR bytes / time / peak.
"""
import json
import os

from . import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIND_NAME = {capi.KIND_SCALAR: "csr_scalar_kernel", capi.KIND_STREAM: "csr_stream_kernel",
             capi.KIND_VECTOR: "csr_vector_kernel", capi.KIND_DICT: "csr_dict_kernel",
             capi.KIND_PATTERN: "csr_pattern_kernel"}
FORMAT_NAME = {capi.KIND_SCALAR: "csr", capi.KIND_STREAM: "csr", capi.KIND_VECTOR: "csr",
               capi.KIND_DICT: "csr-dict16 (lossless: 16-bit value/offset codes, 2 B/nnz)",
               capi.KIND_PATTERN: "csr-pattern8 (lossless: 1-byte row pattern id, 1 B/row)"}
# vector bytes per row of each operation: (reads + writes) of the n-vectors, the diagonal array counted separately
VEC_BYTES = {"spmv": 16, "residual": 24, "jacobi": 24, "spmv_dot": 16, "restrict": 0, "prolong": 0}


def workload(grid, nlevels):
    return (f"AMG-PCG, 3D 7-point Poisson {grid}^3, HEM hierarchy ({nlevels} levels), V(7,7) Jacobi, rel tol 1e-8, "
            f"b=1, x0=0")


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def kernel_label(M, op):
    """the kernel instantiation the library launches for `op` on M (asked from the library itself)"""
    return M.kernel_name({"restrict": "spmv"}.get(op, op))


def csr_bytes(op, nrow, ncol, nnz):
    """algorithmic bytes of `op` on plain CSR"""
    mat = 12 * nnz + 4 * (nrow + 1)
    if op == "jacobi":
        return mat + 32 * nrow                      # x, b, diag in; x' out
    if op == "residual":
        return mat + 8 * ncol + 16 * nrow
    if op in ("spmv", "spmv_dot"):
        return mat + 8 * ncol + 8 * nrow
    if op == "restrict":
        return mat + 8 * ncol + 8 * nrow            # R is nrow(coarse) x ncol(fine)
    if op == "prolong":
        return mat + 8 * ncol + 16 * nrow           # x_f read + written, x_c read
    raise KeyError(op)


def stored_bytes(kind, op, nrow, ncol, nnz):
    """bytes the kernel family `kind` has to move for `op` (see the module docstring)"""
    full = csr_bytes(op, nrow, ncol, nnz)
    if kind == capi.KIND_DICT:
        return full - 10 * nnz
    if kind == capi.KIND_PATTERN:
        vec = full - 12 * nnz - 4 * (nrow + 1)
        if op == "jacobi":
            vec -= 8 * nrow                         # the diagonal comes from the pattern table
        return nrow + vec
    return full


def timed(torch, stream, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def jacobi_traffic(grid, kind):
    """DRAM read+write bytes per launch of the level-0 Jacobi kernel from the committed `ncu --set full` captures"""
    tf = os.path.join(ROOT, "profiles", "jacobi_dram_traffic.json")
    if not os.path.exists(tf):
        return None
    key = {capi.KIND_PATTERN: f"{grid}_pattern", capi.KIND_DICT: f"{grid}_dict"}.get(kind, str(grid))
    return json.load(open(tf)).get(key)


def roofline_jacobi(torch, stream, lib, A0, db, dx, dt, grid, reps=20, where="level 0"):
    """roofline object of the dominant kernel (fused Jacobi sweep) for the kernel family A0 currently selects"""
    kind, tl, _ = A0.kernel()
    n, m, z = A0.nrow, A0.ncol, A0.nnz
    ck = capi.check
    sec = timed(torch, stream, lambda: ck(lib.sparsh_jacobi(A0.h, db.ptr, dx.ptr, dt.ptr, 0.66667, reps)), 1, warm=1) / reps
    peak, peak_kind = measured_peak()
    stored, alg = stored_bytes(kind, "jacobi", n, m, z), csr_bytes("jacobi", n, m, z)
    traffic = jacobi_traffic(grid, kind) if where == "level 0" else None
    out = {"bound": "hbm", "kernel": f"{A0.kernel_name('jacobi')} (fused Jacobi sweep, {where})",
           "achieved": stored / sec / 1e9, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
           "frac": stored / sec / 1e9 / peak, "traffic": traffic, "bytes_per_launch": stored,
           "ms_per_launch": sec * 1e3, "frac_of_8TBs_nominal": stored / sec / 1e9 / 8000.0,
           "format": FORMAT_NAME[kind], "algorithmic_csr_bytes_per_launch": alg,
           "speedup_vs_csr_bound": alg / sec / 1e9 / peak,
           "note": "achieved/frac count the bytes the kernel has to move in the format it reads (stored bytes); "
                   "speedup_vs_csr_bound = plain-CSR algorithmic bytes (SURVEY 8d) / time / peak shows what the lossless "
                   "re-encoding buys over the CSR roofline"}
    if traffic:
        out["traffic_frac_of_peak"] = traffic / sec / 1e9 / peak
    return out


def kernel_table(torch, stream, lib, dH, db, dx, grid, reps=20):
    """per-kernel device time and GB/s of the kernels the solve runs on the finest level (defaults, as selected at upload)
    plus the coarse GEMV: north_star asks for the HBM figure of each kernel, not only of the dominant one"""
    from .device import DeviceVector

    ck = capi.check
    peak, _ = measured_peak()
    A0, P0, R0 = dH.level(0)
    n, z = A0.nrow, A0.nnz
    y, t = DeviceVector(n).fill(0.0), DeviceVector(n).fill(0.0)
    xc, bc = DeviceVector(P0.ncol).fill(0.5), DeviceVector(P0.ncol).fill(0.0)
    dsc = DeviceVector(8).fill(0.0)
    rows = []

    def add(op, label, sec, stored, alg):
        rows.append({"op": op, "kernel": label, "ms": sec * 1e3, "bytes": stored, "gbs": stored / sec / 1e9,
                     "frac": stored / sec / 1e9 / peak, "csr_bytes": alg, "csr_gbs": alg / sec / 1e9})

    def mat_op(op, M, epi, fn, per_call=1):
        kind = M.kernel()[0]
        sec = timed(torch, stream, fn, reps) / per_call
        add(op, kernel_label(M, op), sec, stored_bytes(kind, op, M.nrow, M.ncol, M.nnz), csr_bytes(op, M.nrow, M.ncol, M.nnz))

    mat_op("spmv", A0, "EPI_SPMV", lambda: ck(lib.sparsh_spmv(A0.h, dx.ptr, y.ptr)))
    mat_op("spmv_dot", A0, "EPI_SPMV_DOT", lambda: ck(lib.sparsh_spmv_dot(A0.h, dx.ptr, y.ptr, dsc.ptr)))
    mat_op("residual", A0, "EPI_RESID", lambda: ck(lib.sparsh_residual(A0.h, db.ptr, dx.ptr, y.ptr)))
    mat_op("jacobi", A0, "EPI_JACOBI", lambda: ck(lib.sparsh_jacobi(A0.h, db.ptr, dx.ptr, t.ptr, 0.66667, 2)), per_call=2)
    mat_op("restrict", R0, "EPI_SPMV", lambda: ck(lib.sparsh_restrict(R0.h, y.ptr, bc.ptr)))
    mat_op("prolong", P0, "EPI_PROLONG", lambda: ck(lib.sparsh_prolong_add(P0.h, xc.ptr, y.ptr)))
    for op, label, nb, fn in [
            ("dot", "dot_partial_kernel + reduce_finalize", 16 * n, lambda: ck(lib.sparsh_dot_device(n, dx.ptr, db.ptr, dsc.ptr))),
            ("axpy", "axpy_kernel", 24 * n, lambda: ck(lib.sparsh_axpy(n, 0.5, dx.ptr, y.ptr))),
            ("axpby", "axpby_kernel", 24 * n, lambda: ck(lib.sparsh_axpby(n, 0.5, dx.ptr, 0.25, y.ptr))),
            ("axpbypcz", "axpbypcz_kernel", 32 * n, lambda: ck(lib.sparsh_axpbypcz(n, 0.5, dx.ptr, 0.25, db.ptr, 0.5, y.ptr))),
            ("fill", "fill_kernel", 8 * n, lambda: ck(lib.sparsh_fill(y.ptr, n, 0.0)))]:
        sec = timed(torch, stream, fn, reps)
        add(op, label, sec, nb, nb)
    nl = dH.level(dH.nlevels - 1)[0].nrow
    cb, cx = DeviceVector(nl).fill(1.0), DeviceVector(nl).fill(0.0)
    sec = timed(torch, stream, lambda: ck(lib.sparsh_hierarchy_coarse_solve(dH.h, cb.ptr, cx.ptr)), reps)
    add("coarse_solve", f"dense_gemv_kernel (n_L = {nl}; the inverse is L2-resident, not an HBM figure)", sec,
        8 * nl * nl + 16 * nl, 8 * nl * nl + 16 * nl)
    return rows
