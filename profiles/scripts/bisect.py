"""Option bisect on a small melt (run on the GPU box under a timeout): argv = build_kernel pair_perm pair_pipe [n_side]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import clb_testutil as util
bk, perm, pipe = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 16
m = util.melt(ns, seed=1)
n = len(m["pos"])
v = np.random.default_rng(5).normal(0, 1, (n, 3))
P = util.Pair(m["pos"], m["box"], m["type"], vel=v, state=np.ones(n, np.int32), resid=m["resid"])
for k, val in (("build_kernel", bk), ("pair_perm", perm), ("pair_pipe", pipe)):
    P.e.set_option(k, val)
P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
r, e, f = util.lj_table()
tab = P.add_table(r, e, f, 1)
nb = P.nb_tab(util.type_pairs(2), tab, 2.5)
bl = P.add_list(2, m["bonds"]); al = P.add_list(3, m["angles"])
ib = P.add_bonded(bl); P.bonded_pot(ib, (), "Harmonic", (30.0, 0.97))
ia = P.add_bonded(al); P.bonded_pot(ia, (), "AngularHarmonic", (1.25, np.pi))
a, b = P.e.pairs(), P.o.pairs()
print("pairs equal", len(a) == len(b) and bool((a == b).all()), len(a), flush=True)
P.e.compute_forces(); P.o.compute_forces()
err = util.rel_force_err(P.e.get_particles(fields=("force",))["force"], P.o.get()["force"])
print("force err %.3e" % err, flush=True)
P.both("set_dt", 0.004); P.both("set_langevin", 1, 1.0, 1.0)
P.both("run", 40)
sa, sb = P.e.get_particles(), P.o.get()
dx = np.abs((sa["pos"] + sa["image"] * m["box"]) - (sb["pos"] + sb["image"] * m["box"])).max()
print("traj dx %.3e  build_kernel %d pipe %d" % (dx, P.e.get_option("build_kernel"), P.e.get_option("pair_pipe")), flush=True)
P.close()
print("OK", flush=True)
