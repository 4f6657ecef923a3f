"""Worker of tests/test_dist_plan.py (CPU, gloo, one process per rank): emulates the distributed operators with numpy
from this rank's DistPlan and checks them against the global operators."""
import os
import sys

import numpy as np
import scipy.sparse as sps
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsh_amg_b200 import host  # noqa: E402
from sparsh_amg_b200.distributed import DistPlan  # noqa: E402


def exchange(op, x_local, rank):
    """fill the halo segment exactly as csrc/dist.cu does: packed owned entries out, contiguous halo pieces in"""
    halo = np.zeros(op["nhalo"])
    reqs = []
    sends = []
    for s, q in enumerate(op["send_rank"]):
        buf = torch.from_numpy(np.ascontiguousarray(x_local[op["send_idx"][op["send_ptr"][s]:op["send_ptr"][s + 1]]]))
        sends.append(buf)
        reqs.append(dist.isend(buf, dst=int(q)))
    recvs = []
    for r, q in enumerate(op["recv_rank"]):
        buf = torch.zeros(int(op["recv_ptr"][r + 1] - op["recv_ptr"][r]), dtype=torch.float64)
        recvs.append((r, buf))
        reqs.append(dist.irecv(buf, src=int(q)))
    for q in reqs:
        q.wait()
    for r, buf in recvs:
        halo[op["recv_ptr"][r]:op["recv_ptr"][r + 1]] = buf.numpy()
    return np.concatenate([x_local, halo])


def local_matrix(op):
    return sps.csr_matrix((op["val"], op["colindex"], op["rowptr"]), shape=(op["nrow"], op["ncol_local"] + op["nhalo"]))


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nx, ny, nz = (int(v) for v in sys.argv[1:4])
    coarsening = int(sys.argv[4])
    host.set_options(threads=2, max_levels=32, print_setup=0, coarse_upper=500, coarse_lower=250, coarsening=coarsening)
    A = host.HostMatrix.poisson3d(nx, ny, nz)
    amg = host.HostAmg(A)
    plan = DistPlan(amg, world, rank, tail_threshold=700)
    levels = amg.levels()
    assert plan.nd >= 2, plan.nd
    rng = np.random.default_rng(123)  # same stream on every rank
    for l in range(plan.nd):
        G = levels[l]
        Ag = sps.csr_matrix((G["A"].val, G["A"].colindex, G["A"].rowptr), shape=(G["A"].nrow, G["A"].nrow))
        Pg = sps.csr_matrix((G["P"].val, G["P"].colindex, G["P"].rowptr), shape=(G["P"].nrow, G["P"].ncol))
        rows, rows_c = plan.rows(l), plan.rows(l + 1)
        # ownership is a partition
        gathered = [None] * world
        dist.all_gather_object(gathered, rows.tolist())
        allrows = np.concatenate([np.asarray(g, dtype=np.int64) for g in gathered])
        assert len(allrows) == Ag.shape[0] and len(np.unique(allrows)) == Ag.shape[0]
        assert np.all(np.diff(rows) > 0)
        xg, xc = rng.standard_normal(Ag.shape[0]), rng.standard_normal(Pg.shape[1])
        for which, M, vin, rin, rout in [("A", Ag, xg, rows, rows), ("P", Pg, xc, rows_c, rows),
                                         ("R", Pg.T.tocsr(), xg, rows, rows_c)]:
            op = plan.op(l, which)
            assert op["nrow"] == len(rout) and op["ncol_local"] == len(rin)
            full = exchange(op, vin[rin], rank)
            np.testing.assert_array_equal(full[op["ncol_local"]:], vin[op["halo_global"]])  # halo holds the right entries
            got = local_matrix(op) @ full
            want = (M @ vin)[rout]
            np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-14)
            # interior rows reference no halo entry
            ib, ie = op["interior"]
            seg = op["colindex"][op["rowptr"][ib]:op["rowptr"][ie]]
            assert seg.size == 0 or seg.max() < op["ncol_local"]
            if which == "A":
                np.testing.assert_array_equal(op["diag"], Ag.diagonal()[rows])
                # entry order inside a row is the global one (bit-identical row sums on the device)
                k = len(rows) // 2
                g = rows[k]
                np.testing.assert_array_equal(op["val"][op["rowptr"][k]:op["rowptr"][k + 1]],
                                              G["A"].val[G["A"].rowptr[g]:G["A"].rowptr[g + 1]])
    # a Jacobi sweep + residual + restriction chained through the plans equals the global computation
    G = levels[0]
    Ag = sps.csr_matrix((G["A"].val, G["A"].colindex, G["A"].rowptr), shape=(G["A"].nrow, G["A"].nrow))
    Pg = sps.csr_matrix((G["P"].val, G["P"].colindex, G["P"].rowptr), shape=(G["P"].nrow, G["P"].ncol))
    b = np.ones(Ag.shape[0])
    d = Ag.diagonal()
    x1 = 0.66667 * b / d
    x2 = x1 + 0.66667 * (b - Ag @ x1) / d
    bc = Pg.T @ (b - Ag @ x2)
    rows, rows_c = plan.rows(0), plan.rows(1)
    opA, opR = plan.op(0, "A"), plan.op(0, "R")
    xl = 0.66667 * b[rows] / opA["diag"]
    xl = xl + 0.66667 * (b[rows] - local_matrix(opA) @ exchange(opA, xl, rank)) / opA["diag"]
    rl = b[rows] - local_matrix(opA) @ exchange(opA, xl, rank)
    bcl = local_matrix(opR) @ exchange(opR, rl, rank)
    np.testing.assert_allclose(bcl, bc[rows_c], rtol=1e-13, atol=1e-13)
    dist.barrier()
    if rank == 0:
        print("DIST_PLAN_OK", plan.nd, plan.nlevels)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
