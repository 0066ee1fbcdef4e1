import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import clb_testutil as util
import test_gpu_parity as T
from scipy.spatial import cKDTree
m, P = T._md_pair(n_side=16, langevin=True, dt=0.005, crit=1)
box = np.asarray(m["box"])
def brute(x, r):
    t = cKDTree(np.mod(x, box), boxsize=box)
    return len(t.query_pairs(r))
for k in range(2):
    P.e.run(1)
    st = P.e.get_particles(fields=("pos", "image", "force"))
    print("pos range", st["pos"].min(0), st["pos"].max(0), "image nonzero", int((st["image"] != 0).sum()), "image minmax", st["image"].min(), st["image"].max())
    P.e.energy(0); _, cn = P.e.timers()
    xu = st["pos"] + st["image"] * box
    P.o.set_positions(xu); P.o.compute_forces()
    og = P.o.get()
    print("step %d: pairs<=rc numpy %d  engine %d  oracle %d ; pairs<=rl numpy %d oracle list %d" % (k, brute(st["pos"], 2.5), cn["interacting_pairs"], P.o.count_interacting(), brute(st["pos"], 2.8), len(P.o.pairs())))
    print("   oracle pos after set/rebuild range", og["pos"].min(0), og["pos"].max(0), "max |og.pos mod L - st.pos|", np.abs(np.mod(og["pos"], box) - st["pos"]).max())
    P.o.set_positions(st["pos"]); P.o.compute_forces()
    print("   oracle with folded input: interacting", P.o.count_interacting(), "list", len(P.o.pairs()))
P.close()
