"""Multi-GPU (slab decomposition over NCCL) parity against the oracle: needs >= 2 GPUs in the box, one process each."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nranks", [2, 4])
def test_multi_gpu_matches_oracle(nranks):
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    env = dict(os.environ, MGPU_NSIDE="24" if nranks > 2 else "18")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + nranks), os.path.join(HERE, "mgpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
