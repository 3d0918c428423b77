"""Multi-GPU z-slab parity (needs >= 2 GPUs on the box; skipped otherwise): every rank's slab must equal the
single-domain oracle bit for bit, for the fused pass, the two-sweep kernels, fp32 and PML."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_zslab_ring_bit_exact(world, gpu_count):
    if gpu_count < world:
        pytest.skip(f"needs {world} GPUs, have {gpu_count}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "total mismatches=0" in r.stdout
