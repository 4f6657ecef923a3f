"""CPU coverage of the N>1 path: world_size-2 (and 3) gloo runs of the partition / halo-exchange plans."""
import os
import subprocess
import sys

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
