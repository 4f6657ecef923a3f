"""Worker of tests/test_gpu_dist.py (one process per GPU, NCCL): the N-GPU path against the 1-GPU path on the same
matrix — SpMV bit-identical, V-cycle to 1e-12, PCG residual history to 1e-10 with the same iteration count."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparsh_amg_b200 as sp  # noqa: E402
from sparsh_amg_b200 import host  # noqa: E402
from sparsh_amg_b200.distributed import DistPlan, init_comm  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    threshold = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    use_graph = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    halo_mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group(backend="nccl", device_id=device)
    sp.init(local_rank)
    init_comm(dist, rank, world, device)
    host.set_options(threads=4, max_levels=32, print_setup=0, print_solve=0, coarsening=0, sweeps=7, use_graph=use_graph,
                     halo_mode=halo_mode)
    A = host.HostMatrix.poisson3d(grid, grid, grid)
    amg = host.HostAmg(A)
    n = A.nrow
    # 1-GPU reference on every rank (same arithmetic, single device)
    dH1 = amg.upload()
    rng = np.random.default_rng(7)
    xg = rng.standard_normal(n)
    b = np.ones(n)
    tol = 1e-8 * np.sqrt(n)
    A0, _, _ = dH1.level(0)
    y1 = A0.spmv(sp.DeviceVector(data=xg)).download()
    z1 = dH1.vcycle(sp.DeviceVector(data=b), sp.DeviceVector(n).fill(0.0), 1, x_is_zero=True).download()
    dx1 = sp.DeviceVector(n).fill(0.0)
    it1, hist1, ok1 = dH1.pcg(sp.DeviceVector(data=b), dx1, tol, 500)
    x1 = dx1.download()
    # N-GPU
    plan = DistPlan(amg, world, rank, tail_threshold=threshold)
    dH = plan.upload()
    rows = plan.rows(0)
    assert dH.local_rows(0) == len(rows)
    yl = dH.spmv(0, sp.DeviceVector(data=xg[rows])).download()
    np.testing.assert_array_equal(yl, y1[rows])  # row sums keep their order: bit-identical
    zl = dH.vcycle(sp.DeviceVector(data=b[rows]), sp.DeviceVector(len(rows)).fill(0.0), 1, x_is_zero=True).download()
    np.testing.assert_allclose(zl, z1[rows], rtol=1e-12, atol=1e-12 * np.abs(z1).max())
    dxl = sp.DeviceVector(len(rows)).fill(0.0)
    it, hist, ok = dH.pcg(sp.DeviceVector(data=b[rows]), dxl, tol, 500)
    assert ok and ok1 and abs(it - it1) <= 1, (it, it1)
    m = min(len(hist), len(hist1))
    np.testing.assert_allclose(hist[:m], hist1[:m], rtol=1e-10, atol=1e-13 * hist1[0])
    np.testing.assert_allclose(dxl.download(), x1[rows], rtol=1e-8, atol=1e-10 * np.abs(x1).max())
    # second solve replays the captured graphs and must reproduce the first bit for bit
    it2, hist2, _ = dH.pcg(sp.DeviceVector(data=b[rows]), dxl.fill(0.0), tol, 500)
    assert it2 == it and np.array_equal(hist2, hist)
    # the other entry points on the same handle: AMG as a solver and AMG-preconditioned BiCGStab against their
    # single-GPU twins (same arithmetic per row; only the dot products / norms see another summation tree)
    dxa = sp.DeviceVector(n).fill(0.0)
    ita1, hista1, oka1 = dH1.amg_solve(sp.DeviceVector(data=b), dxa, tol, 200)
    ita, hista, oka = dH.amg_solve(sp.DeviceVector(data=b[rows]), dxl.fill(0.0), tol, 200)
    assert oka and oka1 and ita == ita1, (ita, ita1)
    np.testing.assert_allclose(hista, hista1, rtol=1e-10, atol=1e-13 * hista1[0])
    np.testing.assert_allclose(dxl.download(), dxa.download()[rows], rtol=1e-8, atol=1e-10 * np.abs(x1).max())
    dxb = sp.DeviceVector(n).fill(0.0)
    itb1, histb1, okb1 = dH1.pbicgstab(sp.DeviceVector(data=b), dxb, tol, 500)
    itb, histb, okb = dH.pbicgstab(sp.DeviceVector(data=b[rows]), dxl.fill(0.0), tol, 500)
    assert okb and okb1 and abs(itb - itb1) <= 1, (itb, itb1)
    mb = min(len(histb), len(histb1))
    np.testing.assert_allclose(histb[:mb], histb1[:mb], rtol=1e-7, atol=1e-12 * histb1[0])  # BiCGStab amplifies rounding
    # and PCG once more after the other solvers' graphs were captured: still the same bits
    it3, hist3, _ = dH.pcg(sp.DeviceVector(data=b[rows]), dxl.fill(0.0), tol, 500)
    assert it3 == it and np.array_equal(hist3, hist)
    dist.barrier()
    if rank == 0:
        print(f"DIST_GPU_OK world={world} grid={grid} nd={plan.nd}/{plan.nlevels} pcg={it} (1-GPU {it1}) amg={ita} ({ita1}) "
              f"bicgstab={itb} ({itb1}) halo_mode={halo_mode} graph={use_graph}")
    from sparsh_amg_b200.distributed import shutdown

    shutdown(dist, plan)


if __name__ == "__main__":
    main()
