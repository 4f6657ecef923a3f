"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): N-GPU results against the 1-GPU path."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,grid,threshold,graph,halo_mode,pattern",
                         [(2, 64, 20000, 1, 1, 0), (2, 48, 3000, 0, 1, 0), (2, 64, 20000, 1, 0, 0), (4, 64, 20000, 1, 1, 0),
                          (2, 64, 20000, 1, 1, 1), (2, 96, 30000, 1, 1, 0), (8, 96, 30000, 1, 1, 0)])
def test_ngpu_matches_1gpu(world, grid, threshold, graph, halo_mode, pattern):
    """pattern = 1 selects the csr-pattern8 kernels on every rank (their multi-GPU variants: handshake + fused push)"""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29800 + world + grid % 50),
           os.path.join(HERE, "dist_gpu_worker.py"), str(grid), str(threshold), str(graph), str(halo_mode)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, SPARSH_PATTERN=str(pattern)))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_GPU_OK" in out.stdout
