"""Multi-GPU parity worker: run under torchrun with N >= 2 ranks (one per GPU).

Every rank builds the SAME system, joins the slab decomposition, and checks the N-rank engine against the
fp64 oracle: Verlet pair set (bit-exact), forces (1e-6), energies (1e-8), an MD trajectory with migration and
ghost rebuilds, and a reactive run whose bond list / types / states must equal the oracle's bit-exactly.
Launched by tests/test_gpu_multi.py; prints MGPU_OK on rank 0 when everything passed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import clb_testutil as util

    def say(msg):
        if os.environ.get("MGPU_VERBOSE"):
            print("[rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)
    n_side = int(os.environ.get("MGPU_NSIDE", "18"))
    m = util.melt(n_side, seed=3)
    n = len(m["pos"])
    v = np.random.default_rng(5).normal(0, 1, (n, 3))
    state = np.where(m["type"] == 0, 1, 0).astype(np.int32)
    P = util.Pair(m["pos"], m["box"], m["type"], vel=v, state=state, resid=m["resid"], seed=21, join=True, device=local)
    say("engine joined")
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    r, e, f = util.lj_table()
    tab = P.add_table(r, e, f, 1)
    rl = P.add_list(2, np.zeros((0, 2), np.int64))
    irl = P.add_bonded(rl); P.bonded_pot(irl, (), "Harmonic", (30.0, 0.97))
    nb = P.nb_tab(util.type_pairs(3), tab, 2.5)
    bl = P.add_list(2, m["bonds"]); al = P.add_list(3, m["angles"])
    ib = P.add_bonded(bl); P.bonded_pot(ib, (), "Harmonic", (30.0, 0.97))
    ia = P.add_bonded(al); P.bonded_pot(ia, (), "AngularHarmonic", (1.25, np.pi))
    P.both("exclusions_observe", rl); P.both("exclusions_observe", al)
    P.both("topology_observe", bl); P.both("topology_observe", rl)
    for types in ((1, 0, 0), (1, 0, 2), (0, 2, 1), (2, 0, 1)):
        P.both("topology_register", al, types)
    P.both("topology_initialize")
    say("setup done")
    # 1. pair set
    a, b = P.e.pairs(), P.o.pairs()
    assert len(a) == len(b) and (a == b).all(), "pair set mismatch %d vs %d" % (len(a), len(b))
    say("pairs ok")
    # 2. forces / energies
    P.e.compute_forces(); P.o.compute_forces()
    err = util.rel_force_err(P.e.get_particles(fields=("force",))["force"], P.o.get()["force"])
    assert err < 1e-6, "force mismatch %g" % err
    for k in (nb, ib, ia):
        ea, eb = P.e.energy(k), P.o.energy(k)
        assert abs(ea - eb) <= 1e-8 * abs(eb), ("energy", k, ea, eb)
    say("forces ok")
    # 3. MD with migration + ghost rebuilds
    P.both("set_dt", 0.004); P.both("set_langevin", 1, 1.0, 1.0)
    P.both("run", 60)
    sa, sb = P.e.get_particles(), P.o.get()
    dx = np.abs((sa["pos"] + sa["image"] * m["box"]) - (sb["pos"] + sb["image"] * m["box"])).max()
    assert dx < 5e-4, "trajectory mismatch %g" % dx
    t, c = P.e.timers()
    assert c["rebuilds"] >= 2 and c["ghosts"] > 0, c
    say("md ok")
    # 4. reactions: identical coordinates on both sides, p = 1, nearest partner -> bit-exact topology
    P.o.set_positions(sa["pos"]); P.o.set_velocities(sa["vel"])
    P.both("reaction_general", 1, 10, 1, 0)
    ra = P.e.add_reaction(0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.2, rl, intramolecular=0, intraresidual=0)
    rb = P.o.add_reaction(0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.2, rl, intramolecular=0, intraresidual=0)
    assert ra == rb
    for args in ((ra, 2, 0, 0, 2), (ra, 3, 1, 1, 1)):
        P.e.reaction_add_change(*args); P.o.reaction_add_change(*args)
    na, nbv = P.e.react_now(), P.o.react()
    ca, da = P.e.last_candidates(); cb, db = P.o.candidates()
    assert len(ca) == len(cb) > 20 and (ca == cb).all(), "candidate set mismatch"
    assert na == nbv > 10, ("events", na, nbv)

    def srt(x):
        x = np.asarray(x, np.int64).reshape(len(x), -1)
        return x[np.lexsort(x.T[::-1])] if len(x) else x

    def canon(x):
        x = np.asarray(x, np.int64).copy()
        if len(x):
            fl = x[:, 0] > x[:, -1]; x[fl] = x[fl, ::-1]
        return srt(x)
    assert (srt(P.e.list_get(rl, 2)) == srt(P.o.list_get(rl, 2))).all(), "bond list mismatch"
    assert (canon(P.e.list_get(al, 3)) == canon(P.o.list_get(al, 3))).all(), "angle list mismatch"
    assert P.e.list_size(al) > len(m["angles"])
    ga, gb = P.e.get_particles(fields=("type", "state")), P.o.get()
    assert (ga["type"] == gb["type"]).all() and (ga["state"] == gb["state"]).all(), "type/state mismatch"
    assert (srt(P.e.get_exclusions()) == srt(P.o.get_exclusions())).all()
    say("reaction pass ok")
    # 5. reactive MD: three passes while particles migrate
    P.both("run", 30)
    assert P.e.list_size(rl) == P.o.list_size(rl)
    assert (srt(P.e.list_get(rl, 2)) == srt(P.o.list_get(rl, 2))).all(), "bond list mismatch after reactive run"
    ga, gb = P.e.get_particles(), P.o.get()
    assert (ga["type"] == gb["type"]).all() and (ga["state"] == gb["state"]).all()
    d = ga["pos"] - gb["pos"]; d -= m["box"] * np.rint(d / m["box"])
    dx = np.abs(d).max()
    assert dx < 5e-4, "reactive trajectory mismatch %g" % dx
    ek_a, ek_b = P.e.kinetics()[0], P.o.kinetics()[0]
    assert abs(ek_a - ek_b) < 1e-4 * abs(ek_b), (ek_a, ek_b)
    dist.barrier()
    if rank == 0:
        print("MGPU_OK ranks=%d n=%d pairs=%d force_err=%.2e events=%d bonds=%d ghosts=%d" % (world, n, len(a), err, na, P.e.list_size(rl), c["ghosts"]))
    P.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
