"""Partitioned (multi-GPU) path vs the single-GPU path; needs >= 2 B200s on the box (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from amaru_jl_b200 import lib as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape,n,p2p", [("HEX20", 6, "0"), ("TET10", 5, "0"), ("HEX8", 8, "0"), ("HEX20", 6, "1"), ("TET10", 5, "1"),
                                         ("HEX20", 8, "patch"), ("HEX8", 9, "unfused")])
def test_partitioned_matches_single_gpu(shape, n, p2p):
    """p2p = "0": NCCL halo exchange and all-reduces; "1": peer memory (cudaIpc) with the exchanges fused into the CG kernels (the
    colour-ordered operator on meshes this small); "patch": the same with the patch form of the operator forced; "unfused":
    peer memory with the stand-alone exchange kernels."""
    ngpu = L.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py"), shape, str(n)]
    env = dict(os.environ, AMARU_P2P="0" if p2p == "0" else "1")
    if p2p == "patch":
        env.update(AMARU_EBE_PATCH_MINFILL="0", AMARU_EBE_PATCH_MINPATCH="0")
    if p2p == "unfused":
        env.update(AMARU_P2P_FUSED="0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
