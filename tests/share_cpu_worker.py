"""Worker of tests/test_dist_plan.py::test_shared_hierarchy_under_gloo (CPU, gloo, one process per rank): the
hierarchy-acquisition step of bench.py --gpus N --share-hierarchy, without a GPU."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsh_amg_b200 import host  # noqa: E402
from sparsh_amg_b200.distributed import DistPlan, host_hierarchy, level0_times  # noqa: E402


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    grid = int(sys.argv[1])
    host.set_options(threads=2, max_levels=32, print_setup=0, coarse_upper=500, coarse_lower=250)
    A, amg, shm = host_hierarchy(grid, rank, 2, True, dist.barrier)
    assert A is None and shm is not None and os.path.isdir(shm)
    dims = [amg.level_dims(k) for k in range(amg.nlevels)]
    gathered = [None] * world
    dist.all_gather_object(gathered, dims)
    assert all(g == gathered[0] for g in gathered) and dims[0][0] == grid ** 3
    # every rank can cut its plan out of the mapped copy, and rank 0 can still form A x for the residual check
    plan = DistPlan(amg, world, rank, tail_threshold=700)
    rows = plan.rows(0)
    counts = [None] * world
    dist.all_gather_object(counts, len(rows))
    assert sum(counts) == grid ** 3
    x = np.random.default_rng(7).standard_normal(grid ** 3)
    y = level0_times(A, amg, x)
    ref = host.HostMatrix.poisson3d(grid, grid, grid)
    np.testing.assert_array_equal(y, ref.times(x))
    ref.free()
    plan.free()
    dist.barrier()
    if rank == 0:
        import shutil

        shutil.rmtree(shm, ignore_errors=True)
    print(f"rank {rank} ok", flush=True)


if __name__ == "__main__":
    main()
