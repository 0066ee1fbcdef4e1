"""world_size-2 gloo worker (CPU): host-side logic of the multi-GPU path.

1. the NCCL id hand-off of Engine.join (rank 0 creates, torch.distributed broadcasts) gives every rank the same 128 bytes;
2. the slab decomposition rule (engine.slab_planes / slab_owner == csrc/engine_comm.inl) restated with numpy: every rank
   takes the particles of its own cell planes plus one ghost plane on either side, builds the pairs (i owned, j stored,
   id_i < id_j) inside rc+skin, and the all-gathered union must equal the oracle's global Verlet pair set exactly once
   each (U16: a pair is emitted by the rank owning its lower id)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from chemlab_b200.engine import broadcast_nccl_id, slab_planes
    nid = broadcast_nccl_id()
    got = [None] * world
    dist.all_gather_object(got, nid)
    assert len(nid) == 128 and all(g == nid for g in got) and any(b != 0 for b in nid)

    import clb_testutil as util
    from oracle import pyoracle
    rc, skin = 2.5, 0.3
    m = util.melt(20, seed=9)
    n = len(m["pos"]); box = m["box"]
    L = box[2]
    # the engine stores positions on a 2^32 lattice; use the same quantisation so that plane assignment is identical
    lat = np.rint((m["pos"] / box % 1.0) * 4294967296.0) % 4294967296.0
    pos = lat * (box / 4294967296.0)
    ncz = int(np.floor(L / (rc + skin)))
    planes = slab_planes(ncz, world)
    assert sum(c for _, c in planes) == ncz and planes[0][0] == 0
    assert all(planes[r][0] + planes[r][1] == planes[r + 1][0] for r in range(world - 1))
    cz = np.minimum((lat[:, 2] * ncz / 4294967296.0).astype(np.int64), ncz - 1)
    cz0, nczl = planes[rank]
    local = (cz - cz0) % ncz
    owned = np.where(local < nczl)[0]
    ghost = np.where((local == nczl) | (local == ncz - 1))[0]
    stored = np.concatenate([owned, ghost])
    assert ncz - nczl >= 2                      # the same plane must not be both the upper and the lower ghost
    rl2 = (rc + skin) ** 2
    mine = []
    ps = pos[stored]
    for i in owned:
        d = ps - pos[i]
        d -= box * np.rint(d / box)
        r2 = (d * d).sum(1)
        js = stored[(r2 <= rl2) & (stored > i)]
        mine.append(np.stack([np.full(len(js), i), js], 1))
    mine = np.concatenate(mine)
    allp = [None] * world
    dist.all_gather_object(allp, mine)
    union = np.concatenate(allp)
    union = union[np.lexsort((union[:, 1], union[:, 0]))]
    o = pyoracle.Oracle(n, box, rc, skin, seed=1)
    o.set_particles(pos, np.zeros((n, 3)), np.ones(n), None, m["type"], None, m["resid"])
    o.rebuild()
    ref = o.pairs()
    assert len(union) == len(ref) and (union == ref).all(), (len(union), len(ref))
    owners = [None] * world
    dist.all_gather_object(owners, len(owned))
    assert sum(owners) == n
    dist.barrier()
    if rank == 0:
        print("GLOO_OK world=%d pairs=%d" % (world, len(ref)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
