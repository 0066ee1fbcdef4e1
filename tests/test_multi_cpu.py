"""N > 1 host-side logic on CPU: world_size-2 gloo (tests/gloo_worker.py) + the slab partition rule."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_slab_planes_partition():
    from chemlab_b200.engine import slab_planes, slab_owner
    for ncz in (3, 9, 37, 87):
        for nr in (1, 2, 3, 4, 8):
            if nr > ncz:
                continue
            p = slab_planes(ncz, nr)
            assert p[0][0] == 0 and sum(c for _, c in p) == ncz
            assert max(c for _, c in p) - min(c for _, c in p) <= 1
            assert all(p[r][0] + p[r][1] == p[r + 1][0] for r in range(nr - 1))
    # 1M-bead melt of the bench: 37 planes over 8 ranks -> 5,5,5,5,5,4,4,4
    assert [c for _, c in slab_planes(37, 8)] == [5, 5, 5, 5, 5, 4, 4, 4]
    z = np.array([0.0, 2.79, 2.81, 105.7, -0.1, 105.9])
    own = slab_owner(z, 105.808, 2.8, 8)
    assert own[0] == 0 and own[1] == 0 and own[3] == 7 and own[4] == 7 and own[5] == 0


def test_gloo_world2_decomposition_matches_oracle():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_gloo_world2_driver_host_logic(tmp_path):
    """The driver under torchrun with two ranks (tests/gloo_driver_worker.py): common seed, identical end state, rank 0 alone writes."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(HERE, "gloo_driver_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0 and "GLOO_DRIVER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
