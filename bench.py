#!/usr/bin/env python
"""bench.py — AMG-PCG solve of 3D 7-point Poisson 256^3 (BASELINE.json configs[2]) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU path on the box's host cores

One "step" = one complete AMG-preconditioned CG solve (HEM hierarchy as shipped, V(7,7) Jacobi, omega = 0.66667) of the
16.8M-row system to a RELATIVE residual of 1e-8 (north_star; the reference's absolute 1e-8 is unreachable at this size,
SURVEY F6), x0 = 0, b = 1.  The hierarchy is built once by the host setup and uploaded once (outside the timed region,
as north_star specifies); every step re-solves from x0 = 0.

  value  solve seconds, device-timed (CUDA events on the library's stream), b and x resident in HBM
  e2e    the same solve through the reference-facing host-buffer call (sparsh_hierarchy_solve_host, the twin of
         AMG_GPU1_solver::helper / Solver_PCG_4's cudaMemcpy prologue+epilogue): pinned host b, x -> H2D, solve, D2H of x
  roofline      the dominant kernel (fused Jacobi sweep on the finest level: 14 of the 17 matrix passes per level),
                timed alone with CUDA events on the same stream right after the timed steps; algorithmic bytes from
                SURVEY §8d; peak = MEASURED_PEAKS.json hbm_gbs (burst figure)
  cpu_baseline  the reference's own host code (oracle/_ref: its sources compiled unmodified against the OpenMP MKL shim)
                on all host cores, a bounded sample of the same workload

Inputs are far larger than L2 (the finest matrix alone is 1.4 GB against 126 MB), so no explicit L2 flush is needed.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# PCG iterations to rel 1e-8 on 3D Poisson n^3, b = 1, HEM hierarchy, V(7,7): measured by this repo's solver on a B200
# (profiles/); by parity the reference needs the same count (+-1).  Used ONLY to extrapolate bounded CPU samples.
EXPECTED_PCG_ITERS = {}
_iters_file = os.path.join(ROOT, "profiles", "pcg_iterations.json")
if os.path.exists(_iters_file):
    EXPECTED_PCG_ITERS = {int(k): int(v) for k, v in json.load(open(_iters_file)).items()}

METRIC = "amg_pcg_solve_seconds_poisson3d_256"
UNIT = "s"


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def problem_rhs(n):
    return np.ones(n)


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(grid, iters_per_sample, samples, warm, threads):
    """returns (seconds per PCG iteration, setup seconds, history) using oracle/_ref (the reference's own sources)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bindings import CSR, Oracle, Ref, RefAmg, have_ref

    if not have_ref():
        raise RuntimeError("oracle/_ref/libsparsh_ref.so missing (built where /root/reference exists)")
    o = Oracle.get()
    o.set_threads(threads)
    r = Ref.get()
    r.set_threads(threads)
    A = o.gen_poisson3d(grid, grid, grid)
    ref = RefAmg(A)  # AMG_solver::AMG_solver_setup_jacobi, unmodified, HEM as shipped
    b = problem_rhs(A.nrow)
    times = []
    hist = None
    for s in range(warm + samples):
        t, _, hist = ref.pcg_sample(b, np.zeros(A.nrow), iters_per_sample)
        if s >= warm:
            times.append(t / iters_per_sample)
    return float(np.mean(times)), ref.setup_seconds, hist, ref.nlevels


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    threads = os.cpu_count() or 1
    grid = args.grid
    per_iter, setup_s, hist, nlev = cpu_reference_sample(grid, args.ref_iters, args.steps, min(args.warmup, 1), threads)
    iters = EXPECTED_PCG_ITERS.get(grid)
    sample = (f"{args.steps} samples x {args.ref_iters} PCG iterations of the full {grid}^3 system (reference "
              f"Solver_PCG_1 loop body + its own AMG_solve_jacobi V(7,7) cycle, unmodified sources, OpenMP MKL shim), "
              f"{threads} threads")
    if iters:
        value = per_iter * iters
        sample += f"; solve seconds = seconds/iteration x {iters} iterations (the count this config needs to rel 1e-8)"
    else:
        value = per_iter
        sample += "; iteration count to rel 1e-8 unknown here: value is seconds per PCG ITERATION"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"AMG-PCG, 3D 7-point Poisson {grid}^3, HEM hierarchy ({nlev} levels), V(7,7) "
                                   f"Jacobi, rel tol 1e-8, b=1, x0=0", "grid": grid,
                       "seconds_per_pcg_iteration": per_iter, "reference_setup_seconds": setup_s},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    import sparsh_amg_b200 as sp
    from sparsh_amg_b200 import host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from sparsh_amg_b200 import distributed as dist_mod

        return dist_mod.bench_main(args, METRIC, UNIT, ClockSampler, measured_peak, EXPECTED_PCG_ITERS)

    torch.cuda.set_device(local_rank)
    sp.init(local_rank)
    stream = torch.cuda.Stream()
    sp.set_stream(stream.cuda_stream)
    lib = sp.capi.load()
    grid = args.grid
    threads = os.cpu_count() or 1
    host.set_options(threads=threads, max_levels=32, print_setup=0, print_solve=0, coarsening=0, sweeps=7,
                     use_graph=1)
    t0 = time.time()
    A = host.HostMatrix.poisson3d(grid, grid, grid)
    t_gen = time.time() - t0
    amg = host.HostAmg(A)  # native host setup: HEM + Galerkin, bit-identical integers to the reference's
    rep = host.report()
    dH = amg.upload()       # AMG_GPU1_solver::GPU_Allocations: once
    rep2 = host.report()
    n = A.nrow

    # pinned host buffers for the e2e leg, device-resident vectors for the device-timed leg
    hb, hx = C.c_void_p(), C.c_void_p()
    sp.capi.check(lib.sparsh_host_alloc(n * 8, C.byref(hb)))
    sp.capi.check(lib.sparsh_host_alloc(n * 8, C.byref(hx)))
    b_host = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_double)), shape=(n,))
    x_host = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_double)), shape=(n,))
    b_host[:] = problem_rhs(n)
    tol = 1e-8 * float(np.linalg.norm(b_host))
    db, dx = sp.DeviceVector(data=b_host), sp.DeviceVector(n)
    max_iter = args.max_iter

    def solve_device():
        dx.fill(0.0)
        return dH.pcg(db, dx, tol, max_iter)

    def solve_host():
        x_host[:] = 0.0
        return dH.solve_host("pcg", b_host, x_host, tol, max_iter)

    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    for _ in range(n_warm):
        it, hist, ok = solve_device()
    if not ok and not args.profile:
        raise RuntimeError(f"PCG did not converge in {max_iter} iterations (last residual {hist[-1]:.3e})")

    sampler = ClockSampler(local_rank)
    sampler.start()
    sp.launch_count(reset=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    sp.sync()
    ev[0].record(stream)
    for _ in range(args.steps):
        it, hist, ok = solve_device()
    ev[1].record(stream)
    ev[1].synchronize()
    launches = sp.launch_count()
    solve_s = ev[0].elapsed_time(ev[1]) * 1e-3 / args.steps

    # e2e: host buffers in, host buffer out, copies inside the timed region
    for _ in range(2):
        solve_host()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sp.sync()
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        it_h, hist_h, ok_h = solve_host()
    e1.record(stream)
    e1.synchronize()
    e2e_wall = (time.perf_counter() - w0) / args.steps
    e2e_s = max(e0.elapsed_time(e1) * 1e-3 / args.steps, e2e_wall)  # wall clock covers the host-side orchestration too
    clocks = sampler.stop()
    x_final = x_host.copy()
    if args.dump_hist:  # residual history of the solve, for the pin against the reference's own history (tools/)
        json.dump({"grid": grid, "iterations": int(it), "tol": tol, "history": [float(v) for v in hist]},
                  open(args.dump_hist, "w"))

    # roofline of the dominant kernel: fused Jacobi sweep on the finest level, timed alone on the same stream
    A0, _, _ = dH.level(0)
    z = A0.nnz
    kind, tl, _ = A0.kernel()
    dict_fmt = kind == sp.capi.KIND_DICT
    pat_fmt = kind == sp.capi.KIND_PATTERN  # opt-in (SPARSH_PATTERN=1)
    jac_bytes = 12 * z + 4 * (n + 1) + 32 * n          # ALGORITHMIC bytes of the CSR operation
    jac_stored = (2 if dict_fmt else 12) * z + 4 * (n + 1) + 32 * n  # bytes the kernel actually has to move
    if pat_fmt:
        jac_stored = 25 * n  # pattern byte + b + x + x' per row; the diagonal comes from the pattern table
    tb = sp.DeviceVector(n)
    reps = 20
    lib.sparsh_jacobi(A0.h, db.ptr, dx.ptr, tb.ptr, 0.66667, 4)
    j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    j0.record(stream)
    lib.sparsh_jacobi(A0.h, db.ptr, dx.ptr, tb.ptr, 0.66667, reps)
    j1.record(stream)
    j1.synchronize()
    jac_s = j0.elapsed_time(j1) * 1e-3 / reps
    peak, peak_kind = measured_peak()
    achieved = jac_bytes / jac_s / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "jacobi_dram_traffic.json")
    if os.path.exists(tf):
        traffic = json.load(open(tf)).get(f"{grid}_pattern" if pat_fmt else f"{grid}_dict" if dict_fmt else str(grid))
    vbytes = dH.vcycle_bytes(True)
    iter_bytes = vbytes + (12 * z + 4 * (n + 1) + 16 * n) + 48 * n + 16 * n + 24 * n  # + SpMV, x/r update, dot, p update

    # true residual of the returned solution (host check of the device result, outside any timed region)
    r_true = float(np.linalg.norm(b_host - A.times(x_final)))

    cpu = None
    if not args.no_cpu_baseline:
        try:
            per_iter, setup_s, _, _ = cpu_reference_sample(grid, args.ref_iters, 1, 0, threads)
            cpu = {"value": per_iter * it, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"{args.ref_iters} PCG iterations of the full {grid}^3 system by the reference's own host "
                             f"code (oracle/_ref: unmodified sources + OpenMP MKL shim), seconds/iteration x {it} "
                             f"iterations; reference setup {setup_s:.1f}s not included",
                   "seconds_per_pcg_iteration": per_iter}
        except Exception as e:  # the checker is optional for the measurement itself
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"unavailable: {e}"}

    line = {"metric": METRIC, "value": solve_s, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": solve_s * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"AMG-PCG, 3D 7-point Poisson {grid}^3, HEM hierarchy ({amg.nlevels} levels), "
                                   f"V(7,7) Jacobi, rel tol 1e-8, b=1, x0=0", "grid": grid, "rows": n, "nnz": z,
                       "pcg_iterations": it, "final_rel_residual": float(hist[-1] / hist[0]),
                       "true_rel_residual": r_true / float(np.linalg.norm(b_host)),
                       "ms_per_pcg_iteration": solve_s * 1e3 / max(it, 1),
                       "iteration_algorithmic_gb": iter_bytes / 1e9,
                       "solve_effective_gbs": iter_bytes * it / solve_s / 1e9,
                       "l2": "inputs larger than L2 (finest matrix 1.4 GB vs 126 MB)", "cuda_graph": True,
                       "host_setup_seconds": rep["setup_seconds"], "upload_seconds": rep2["upload_seconds"],
                       "matrix_generation_seconds": t_gen},
            "e2e": {"value": e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": n * 8,
                    "pcg_iterations": it_h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm",
                         "kernel": f"{'csr_pattern_kernel' if pat_fmt else 'csr_dict_kernel' if dict_fmt else 'csr_stream_kernel'}<{tl},EPI_JACOBI> "
                                   "(fused Jacobi sweep, level 0)",
                         "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "bytes_per_launch": jac_bytes,
                         "ms_per_launch": jac_s * 1e3, "frac_of_8TBs_nominal": achieved / 8000.0,
                         "format": "csr-pattern8 (lossless: 1-byte row pattern id)" if pat_fmt else
                                   "csr-dict16 (lossless: 16-bit value/offset codes, 2 B/nnz)" if dict_fmt else "csr",
                         "stored_bytes_per_launch": jac_stored, "stored_gbs": jac_stored / jac_s / 1e9,
                         "stored_frac": jac_stored / jac_s / 1e9 / peak,
                         "note": ("achieved counts the ALGORITHMIC CSR bytes (12 B/nnz); the kernel reads the lossless "
                                  "csr-dict16 twin, so frac > 1 is compression, stored_frac is the DRAM-roofline "
                                  "fraction of what is really moved") if dict_fmt else ""},
            "cpu_baseline": cpu, "clocks": clocks}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="grid side n of the n^3 Poisson problem")
    ap.add_argument("--max-iter", type=int, default=1000)
    ap.add_argument("--ref-iters", type=int, default=2, help="PCG iterations per bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tail-threshold", type=int, default=1100000,
                    help="N>1: levels with at most this many rows are replicated on every GPU instead of partitioned")
    ap.add_argument("--share-hierarchy", action="store_true",
                    help="N>1: rank 0 builds the host hierarchy once, the other ranks map it (needed beyond 256^3)")
    ap.add_argument("--halo-mode", type=int, default=1, help="N>1: 1 NVLink peer-memory pushes, 0 ncclSend/ncclRecv")
    ap.add_argument("--dump-hist", default=None, help="N=1: write the PCG residual history of the solve to this JSON file")
    ap.add_argument("--profile", action="store_true",
                    help="for ncu: honour --warmup/--max-iter literally, do not insist on convergence")
    args = ap.parse_args()
    if args.grid != 256:  # the headline metric is quoted at 256^3; other grids are named for what they are
        global METRIC
        METRIC = f"amg_pcg_solve_seconds_poisson3d_{args.grid}"
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started without the launcher: become `torchrun --nproc-per-node N bench.py ...` (one process per GPU)
        port = str(29500 + os.getpid() % 400)
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                  f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", port,
                                  os.path.abspath(__file__)] + sys.argv[1:])
    run_ours(args)


if __name__ == "__main__":
    main()
