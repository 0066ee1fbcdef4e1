"""world_size-2 gloo worker (CPU): the multi-rank host logic of the chemlab driver -- every rank runs the same driver (here on the
oracle adapter, which simulates the whole system on each rank), the seed drawn on rank 0 reaches every rank, rank 0 alone
prints and writes the products, the ranks end in the same state."""
import hashlib
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import chemlab_b200.espressopp._context as C
    from oracle.engine_adapter import OracleEngine
    from chemlab_b200 import start_simulation as S
    C.Engine = OracleEngine
    base = sys.argv[1]
    d = os.path.join(base, "rank%d" % rank)               # separate working directories: what a rank writes is visible
    shutil.copytree(os.path.join(HERE, "golden", "atrp_lj"), d)
    os.chdir(d)
    # no --rng_seed: rank 0 draws one and broadcasts it (every rank must thermalise and react identically)
    r = S.main(["@params", "--run", "400", "--start_ar", "200", "--energy_collect", "200"])
    seed = int(r["prefix"].rsplit("_", 1)[1])
    seeds = [None] * world
    dist.all_gather_object(seeds, seed)
    assert len(set(seeds)) == 1, seeds
    g = r["system"]._ctx.engine.get_particles(fields=("pos", "type", "state"))
    h = hashlib.sha1(np.ascontiguousarray(g["pos"]).tobytes() + g["type"].tobytes() + g["state"].tobytes()).hexdigest()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    assert len(set(hs)) == 1, "ranks diverged"
    files = sorted(os.listdir("data")) if os.path.isdir("data") else []
    products = [f for f in files if f.endswith(("_confout.gro", "_output_topol.top", "_bonds.dat", "_reaction_counters", "_benchmark.csv", "_topology.dat"))]
    if rank == 0:
        assert len(products) >= 6, files
    else:
        assert not products, products
    dist.barrier()
    if rank == 0:
        print("GLOO_DRIVER_OK")


if __name__ == "__main__":
    main()
