"""CPU coverage of the N>1 path: world_size-2 (and 3) gloo runs of the partition / halo-exchange plans."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,shape,coarsening", [(2, (20, 20, 24), 0), (3, (16, 18, 20), 0), (2, (20, 20, 24), 1),
                                                    (8, (16, 16, 32), 0)])
def test_partition_plans_with_gloo(world, shape, coarsening):
    port = 29500 + (os.getpid() % 2000) + world + 7 * coarsening
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "dist_cpu_worker.py"),
           *[str(s) for s in shape], str(coarsening)]
    env = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_PLAN_OK" in out.stdout


def test_shared_hierarchy_under_gloo():
    """bench.py --gpus N --share-hierarchy, the host part: rank 0 builds and publishes, every rank (0 included) maps the
    same copy, cuts its plan out of it, and rank 0 can still form A x for the residual check (world size 2, gloo)."""
    port = 29500 + (os.getpid() % 2000) + 31
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "share_cpu_worker.py"), "20"]
    env = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="", MASTER_PORT=str(port))
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_local_blocks_keep_the_compressed_formats():
    """The [owned | halo] relabelling of a slab partition shifts all halo columns of a block by one constant, so the
    local blocks of a stencil hierarchy still qualify for csr-dict16 and csr-pattern8 (no escape rows)."""
    from sparsh_amg_b200 import host
    from sparsh_amg_b200.distributed import DistPlan
    from sparsh_amg_b200.generators import HostCSR
    from test_host_setup import _check_pattern_decode, _dict_encode, _pattern_encode

    host.set_options(max_levels=32, print_setup=0)
    A = host.HostMatrix.poisson3d(32, 32, 32)
    amg = host.HostAmg(A)
    for rank in (0, 2):
        p = DistPlan(amg, 4, rank, tail_threshold=3000)
        assert p.nd >= 2
        for l in range(p.nd):
            a = p.op(l, "A")
            M = HostCSR(a["nrow"], a["ncol_local"] + a["nhalo"], a["rowptr"], a["colindex"], a["val"])
            enc = _pattern_encode(M, np.ascontiguousarray(a["diag"]))
            _check_pattern_decode(M, *enc)
            assert 0 < enc[4] <= 27 and enc[5] == 0
            _, dval, doff = _dict_encode(M)
            assert 0 < len(dval) <= 3 and 0 < len(doff) <= 9
        p.free()
    amg.free()
    A.free()
    host.set_options(max_levels=6, print_setup=1)
