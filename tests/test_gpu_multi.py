"""Partitioned (multi-GPU) path vs the single-GPU path; needs >= 2 B200s on the box (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from amaru_jl_b200 import lib as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape,n,p2p", [("HEX20", 6, "0"), ("TET10", 5, "0"), ("HEX8", 8, "0"), ("HEX20", 6, "1"), ("TET10", 5, "1")])
def test_partitioned_matches_single_gpu(shape, n, p2p):
    """p2p = "1": the CG loop's halo exchange and scalar all-reduces run as peer-memory kernels (cudaIpc) instead of NCCL."""
    ngpu = L.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py"), shape, str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, AMARU_P2P=p2p))
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
