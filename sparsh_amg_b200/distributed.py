"""Multi-GPU glue: one process per GPU (torchrun), torch.distributed for the control plane, the library's own NCCL
communicator (csrc/dist.cu) for the data path.

  DistPlan        this rank's part of a host hierarchy (host/dist_plan.cpp): local operators + exchange plans, exposed as
                  numpy arrays so the partition logic is testable on CPU with gloo
  DistHierarchy   the uploaded distributed hierarchy (sparsh_dist_t): spmv / vcycle / pcg on local row blocks
  bench_main      the N>1 leg of bench.py
"""
import ctypes as C
import json
import os
import time

import numpy as np

from . import capi, host
from .capi import check, dp
from .device import DeviceVector

c_int_p = C.POINTER(C.c_int)


def _np(ptr, n, dtype):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,))


class DistPlan:
    def __init__(self, amg, nranks, rank, tail_threshold=1100000):
        self.lib = host.load()
        self.lib.sparsh_host_dist_plan.restype = C.c_void_p
        self.lib.sparsh_host_dist_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.sparsh_host_dist_plan_free.argtypes = [C.c_void_p]
        self.lib.sparsh_host_dist_plan_levels.argtypes = [C.c_void_p, c_int_p, c_int_p]
        self.lib.sparsh_host_dist_plan_rows.argtypes = [C.c_void_p, C.c_int, C.POINTER(c_int_p)]
        self.lib.sparsh_host_dist_plan_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(capi.DistOpDesc),
                                                      C.POINTER(c_int_p)]
        self.lib.sparsh_host_dist_upload.restype = C.c_void_p
        self.lib.sparsh_host_dist_upload.argtypes = [C.c_void_p]
        self.amg = amg
        self.nranks, self.rank = nranks, rank
        self.h = self.lib.sparsh_host_dist_plan(amg.h, nranks, rank, int(tail_threshold))
        if not self.h:
            raise capi.SparshError("hierarchy has a single level: nothing to distribute")
        a, b = C.c_int(), C.c_int()
        self.lib.sparsh_host_dist_plan_levels(self.h, C.byref(a), C.byref(b))
        self.nd, self.nlevels = a.value, b.value

    def rows(self, level):
        """owned global row ids of `level` (0..nd), ascending"""
        p = c_int_p()
        n = self.lib.sparsh_host_dist_plan_rows(self.h, level, C.byref(p))
        return _np(p, n, np.int32)

    def op(self, level, which):
        """which: 'A' | 'P' | 'R' -> dict of numpy views over this rank's local operator and its exchange plan"""
        d = capi.DistOpDesc()
        hg = c_int_p()
        self.lib.sparsh_host_dist_plan_op(self.h, level, "APR".index(which), C.byref(d), C.byref(hg))
        out = dict(nrow=d.nrow, ncol_local=d.ncol_local, nhalo=d.nhalo, nnz=d.nnz,
                   rowptr=_np(d.rowptr, d.nrow + 1, np.int32), colindex=_np(d.colindex, d.nnz, np.int32),
                   val=_np(d.val, d.nnz, np.float64), diag=_np(d.diag, d.nrow, np.float64) if d.diag else None,
                   send_rank=_np(d.send_rank, d.n_send, np.int32), send_ptr=_np(d.send_ptr, d.n_send + 1, np.int32),
                   recv_rank=_np(d.recv_rank, d.n_recv, np.int32), recv_ptr=_np(d.recv_ptr, d.n_recv + 1, np.int32),
                   interior=(d.interior_begin, d.interior_end), halo_global=_np(hg, d.nhalo, np.int32))
        out["send_idx"] = _np(d.send_idx, int(out["send_ptr"][-1]) if d.n_send else 0, np.int32)
        return out

    def upload(self):
        dh = self.lib.sparsh_host_dist_upload(self.h)
        if not dh:
            raise capi.SparshError(capi.load().sparsh_last_error().decode(errors="replace"))
        return DistHierarchy(dh, self)

    def free(self):
        if getattr(self, "h", None):
            self.lib.sparsh_host_dist_plan_free(self.h)
            self.h = None


class DistHierarchy:
    def __init__(self, handle, plan):
        self.lib = capi.load()
        self.h = handle
        self.plan = plan

    def local_rows(self, level=0):
        n = C.c_int()
        check(self.lib.sparsh_dist_local_rows(self.h, level, C.byref(n)))
        return n.value

    def spmv(self, level, x, y=None):
        y = y or DeviceVector(self.local_rows(level))
        check(self.lib.sparsh_dist_spmv(self.h, level, x.ptr, y.ptr))
        return y

    def vcycle(self, b, x, cycles=1, x_is_zero=False):
        check(self.lib.sparsh_dist_vcycle(self.h, b.ptr, x.ptr, cycles, int(bool(x_is_zero))))
        return x

    def level_matrix(self, level):
        """borrowed local block of A_level (columns = [owned | halo])"""
        from .device import DeviceMatrix

        a = C.c_void_p()
        check(self.lib.sparsh_dist_level_matrix(self.h, level, C.byref(a)))
        return DeviceMatrix(handle=a.value, owned=False)

    def _solve(self, fn, b, x, tol, max_iter):
        hist = np.zeros(max_iter + 1)
        it = C.c_int()
        rc = check(fn(self.h, b.ptr, x.ptr, float(tol), int(max_iter), dp(hist), C.byref(it)), allow_not_converged=True)
        return it.value, hist[: it.value + 1], rc == capi.SPARSH_OK

    def pcg(self, b, x, tol, max_iter=1000):
        return self._solve(self.lib.sparsh_dist_pcg, b, x, tol, max_iter)

    def amg_solve(self, b, x, tol, max_cycles=500):
        return self._solve(self.lib.sparsh_dist_amg_solve, b, x, tol, max_cycles)

    def pbicgstab(self, b, x, tol, max_iter=1000):
        return self._solve(self.lib.sparsh_dist_pbicgstab, b, x, tol, max_iter)


def init_comm(torch_dist, rank, world, device):
    """create the library's NCCL communicator; the unique id travels over torch.distributed"""
    import torch

    lib = capi.load()
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(lib.sparsh_dist_get_unique_id(buf))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if torch_dist.get_backend() == "nccl":
        t = t.to(device)
    torch_dist.broadcast(t, src=0)
    raw = bytes(t.cpu().numpy().tobytes())
    check(lib.sparsh_dist_init(raw, world, rank))


def host_hierarchy(grid, rank, threads, share, barrier):
    """(A, amg, shm_dir) for bench_main.  share=False: every rank builds its own copy (A is the level-0 matrix).
    share=True: rank 0 builds the hierarchy once with all cores, publishes it as files and drops its private copy; every
    rank (0 included) maps the files read-only (host/share.cpp) — one copy per node in the page cache; A is None.
    Needed beyond 256^3: at 512^3 eight private copies (8 x 28 GB) do not fit in a box's RAM."""
    if not share:
        A = host.HostMatrix.poisson3d(grid, grid, grid)
        return A, host.HostAmg(A), None  # (sequential) host hierarchy on every rank; only its part goes to the GPU
    base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    shm = os.path.join(base, f"sparsh_hier_{os.environ.get('MASTER_PORT', '0')}_{grid}")
    if rank == 0:
        host.set_options(threads=os.cpu_count() or 1)  # the other ranks are waiting: all cores to the one setup
        A = host.HostMatrix.poisson3d(grid, grid, grid)
        amg = host.HostAmg(A)
        amg.save(shm)
        amg.free()
        A.free()
        host.set_options(threads=threads)
    barrier()
    return None, host.HostAmg.load(shm), shm


def level0_times(A, amg, x):
    """A_0 x on the host (true-residual check outside the timed region); A is None with a shared hierarchy"""
    if A is not None:
        return A.times(x)
    import scipy.sparse as sps

    L0 = amg.levels()[0]["A"]
    return sps.csr_matrix((L0.val, L0.colindex, L0.rowptr), shape=(L0.nrow, L0.nrow)) @ x


def bench_main(args, METRIC, UNIT, ClockSampler):
    """bench.py --gpus N (N > 1), launched by torchrun: strong scaling of the same solve, rows split across ranks"""
    import torch
    import torch.distributed as dist

    import sparsh_amg_b200 as sp

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group(backend="nccl", device_id=device)
    sp.init(local_rank)
    stream = torch.cuda.Stream()
    sp.set_stream(stream.cuda_stream)
    init_comm(dist, rank, world, device)

    grid = args.grid
    threads = max(1, (os.cpu_count() or 1) // world)
    host.set_options(threads=threads, max_levels=32, print_setup=0, print_solve=0, coarsening=0, sweeps=7, use_graph=1,
                     halo_mode=args.halo_mode)
    t0 = time.time()
    A, amg, shm = host_hierarchy(grid, rank, threads, getattr(args, "share_hierarchy", False), dist.barrier)
    t_setup = time.time() - t0
    plan = DistPlan(amg, world, rank, tail_threshold=args.tail_threshold)
    dH = plan.upload()
    if shm is not None:
        dist.barrier()
        if rank == 0:
            import shutil

            shutil.rmtree(shm, ignore_errors=True)
    n_local = dH.local_rows(0)
    n = amg.level_dims(0)[0]
    rows = plan.rows(0)
    b_local = np.ones(n_local)
    tol = 1e-8 * float(np.sqrt(n))
    db, dx = DeviceVector(data=b_local), DeviceVector(n_local)
    max_iter = args.max_iter

    def solve():
        dx.fill(0.0)
        return dH.pcg(db, dx, tol, max_iter)

    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    for _ in range(n_warm):
        it, hist, ok = solve()
    if not ok and not args.profile:
        raise RuntimeError(f"rank {rank}: PCG did not converge ({hist[-1]:.3e})")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sp.launch_count(reset=True)
    dist.barrier()
    sp.sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        it, hist, ok = solve()
    e1.record(stream)
    e1.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    t_local = torch.tensor([e0.elapsed_time(e1) * 1e-3 / args.steps], device=device, dtype=torch.float64)
    dist.all_reduce(t_local, op=dist.ReduceOp.MAX)  # device time, max over ranks
    solve_s = float(t_local.item())
    launches = sp.launch_count()

    # e2e: host buffers in (each rank its row block), host buffers out
    lib = capi.load()
    hb, hx = C.c_void_p(), C.c_void_p()
    check(lib.sparsh_host_alloc(n_local * 8, C.byref(hb)))
    check(lib.sparsh_host_alloc(n_local * 8, C.byref(hx)))
    b_host = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_double)), shape=(n_local,))
    x_host = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_double)), shape=(n_local,))
    b_host[:] = 1.0

    def solve_host():
        check(lib.sparsh_memcpy_h2d(db.ptr, b_host.ctypes.data, n_local * 8))
        check(lib.sparsh_memcpy_h2d(dx.ptr, x_host.ctypes.data, n_local * 8))
        r = dH.pcg(db, dx, tol, max_iter)
        check(lib.sparsh_memcpy_d2h(x_host.ctypes.data, dx.ptr, n_local * 8))
        return r

    # (as at N=1 each step is timed by itself: a rank writing the initial guess x0 = 0 into its host buffer is the
    # caller preparing its input and stays outside, behind a barrier; the copies of b, x0 and x are inside)
    x_host[:] = 0.0
    solve_host()
    t_acc = 0.0
    for _ in range(args.steps):
        x_host[:] = 0.0
        dist.barrier()
        w0 = time.perf_counter()
        solve_host()
        t_acc += time.perf_counter() - w0
    t_e2e = torch.tensor([t_acc / args.steps], device=device, dtype=torch.float64)
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None

    # roofline of the dominant kernel on this rank's row block of the finest level (Jacobi sweep, no exchange), timed
    # alone with CUDA events on the same stream; the line carries the slowest rank's figure
    from . import benchutil

    try:
        A0 = dH.level_matrix(0)
        xa, xb2 = DeviceVector(A0.ncol + 8).fill(0.5), DeviceVector(A0.ncol + 8).fill(0.0)
        roof = benchutil.roofline_jacobi(torch, stream, lib, A0, db, xa, xb2, grid, where="one rank's row block of level 0")
        t_j = torch.tensor([roof["ms_per_launch"]], device=device, dtype=torch.float64)
        dist.all_reduce(t_j, op=dist.ReduceOp.MAX)
        scale = roof["ms_per_launch"] / float(t_j.item())  # rescale rank 0's figures to the slowest rank's time
        for key in ("achieved", "frac", "frac_of_8TBs_nominal", "speedup_vs_csr_bound"):
            roof[key] *= scale
        roof["ms_per_launch"] = float(t_j.item())
        roof["traffic"] = None  # the ncu capture is of the single-GPU launch; a rank's block moves 1/N of it
    except Exception as e:  # the timing above must survive a failure of the reporting extras
        roof = {"bound": "hbm", "error": f"{type(e).__name__}: {e}"}

    # true residual of the assembled solution, checked on rank 0's host (outside every timed region)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=device))
    maxc = int(max(c.item() for c in counts))
    pad = torch.zeros(maxc, dtype=torch.float64, device=device)
    pad[:n_local] = torch.from_numpy(x_host.copy()).to(device)
    padr = torch.full((maxc,), -1, dtype=torch.int64, device=device)
    padr[:n_local] = torch.from_numpy(rows.astype(np.int64)).to(device)
    xs = [torch.zeros_like(pad) for _ in range(world)]
    rs = [torch.zeros_like(padr) for _ in range(world)]
    dist.all_gather(xs, pad)
    dist.all_gather(rs, padr)
    if rank == 0:
        x_full = np.zeros(n)
        for xr, rr in zip(xs, rs):
            rr = rr.cpu().numpy()
            m = rr >= 0
            x_full[rr[m]] = xr.cpu().numpy()[m]
        r_true = float(np.linalg.norm(np.ones(n) - level0_times(A, amg, x_full)))
        line = {"metric": METRIC, "value": solve_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": solve_s * 1e3, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": benchutil.workload(grid, amg.nlevels), "grid": grid, "rows": n,
                           "nnz": int(amg.level_dims(0)[1])},
                "details": {"partition": f"rows split over {world} GPUs, {plan.nd} distributed levels + "
                                         f"{plan.nlevels - plan.nd} replicated",
                            "pcg_iterations": it, "final_rel_residual": float(hist[-1] / hist[0]),
                            "true_rel_residual": r_true / float(np.sqrt(n)),
                            "ms_per_pcg_iteration": solve_s * 1e3 / max(it, 1), "timing": "CUDA events, max over ranks",
                            "l2": "inputs larger than L2 per rank at the finest levels", "cuda_graph": True,
                            "host_setup_seconds": t_setup, "tail_threshold_rows": args.tail_threshold,
                            "halo_exchange": "NVLink peer-memory push + flags" if args.halo_mode == 1 else "ncclSend/ncclRecv"},
                "e2e": {"value": float(t_e2e.item()), "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8,
                        "d2h_bytes_per_step": n * 8},
                "gpu_launches": int(launches) * world,
                "roofline": roof, "cpu_baseline": None, "clocks": clocks}
        print(json.dumps(line), flush=True)
    shutdown(dist, plan)


def shutdown(dist, plan=None):
    """Leave without tearing communicators down.  Destroying an NCCL communicator whose kernels live in instantiated
    CUDA graphs blocked at process exit on the B200 box (both torch's and ours are alive here), so: drain, barrier,
    flush, and let the OS reclaim the process."""
    import sys

    import torch

    capi.load().sparsh_sync()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)
