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


def test_driver_on_two_gpus_equals_one_gpu(tmp_path):
    """`torchrun --nproc-per-node 2 -m chemlab_b200.start_simulation @params` (the mpirun of this engine) must leave the same
    chemistry behind as the single-GPU run of examples/atrp_lj: identical types/states and reaction bonds."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import shutil
    import numpy as np
    root = os.path.dirname(HERE)
    outs = {}
    for tag, launcher in (("one", [sys.executable, "-m", "chemlab_b200.start_simulation"]),
                          ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                                   "--master-port", "29641", "-m", "chemlab_b200.start_simulation"])):
        d = str(tmp_path / tag)
        shutil.copytree(os.path.join(HERE, "golden", "atrp_lj"), d)
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        with open(os.path.join(d, "params"), "a") as f:       # overrides go into the arg-file (torchrun would eat `--run`)
            f.write("\nrng_seed=42\nrun=800\nstart_ar=200\nenergy_collect=200\n")
        r = subprocess.run(launcher + ["@params"], cwd=d, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        outs[tag] = (np.loadtxt(os.path.join(d, "data", "cpc01_42_state.dat")), np.loadtxt(os.path.join(d, "data", "cpc01_42_bonds_chem_0.dat"), ndmin=2),
                     open(os.path.join(d, "data", "cpc01_42_benchmark.csv")).read().split())
    assert (outs["one"][0] == outs["two"][0]).all()
    a, b = outs["one"][1], outs["two"][1]
    assert a.shape == b.shape and len(a) > 0
    assert (a[np.lexsort((a[:, 1], a[:, 0]))] == b[np.lexsort((b[:, 1], b[:, 0]))]).all()
    assert outs["one"][2][0] == "1" and outs["two"][2][0] == "2"          # nranks column of the benchmark record
