import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import clb_testutil as util
from oracle import pyoracle
m = util.melt(12, seed=5); n = len(m["pos"])
state = np.where(m["type"] == 0, 1, 0).astype(np.int32)
o = pyoracle.Oracle(n, m["box"], 2.5, 0.3, seed=99)
o.set_particles(m["pos"], np.zeros((n,3)), np.ones(n), None, m["type"], state, m["resid"])
o.set_exclusions(util.exclusions_from(m["bonds"], m["angles"]))
r,e,f = util.lj_table(); tab = o.add_table(r,e,f,1)
rl = o.add_list(2); irl = o.add_bonded(rl); o.bonded_set_potential(irl, (), 1, (30.0,0.97))
nb = o.add_nonbonded(1)
for a,b in util.type_pairs(3): o.nb_set_tab(nb,a,b,tab,2.5)
bl = o.add_list(2); o.list_add(bl, m["bonds"]); al = o.add_list(3); o.list_add(al, m["angles"]); ql = o.add_list(4)
o.set_dt(0.004); o.reaction_general(1,10,1,0)
o.excl_observe(rl); o.excl_observe(al); o.excl_observe(ql)
o.tm_observe(bl); o.tm_observe(rl)
for lst, types in ((al, (1, 0, 0)), (al, (1, 0, 2)), (al, (0, 2, 2)), (al, (3, 0, 2)), (al, (0, 2, 3)), (ql, (1, 0, 0, 1)), (ql, (0, 1, 0, 0)),
                   (ql, (0, 1, 0, 2)), (ql, (3, 0, 2, 3)), (ql, (4, 3, 0, 2))):
    o.tm_register(lst, types)
o.tm_initialize()
rr = o.add_reaction(0,0,1,0,1,2,1,2,1e6,1.25,rl,intramolecular=1,intraresidual=0)
for side, lvl, old, new, kw in ((2, 0, 0, 2, dict(new_mass=1.5)), (3, 1, 1, 3, dict(state_mode=1, state_value=1)),
                                (3, 2, 1, 3, dict(state_mode=1, state_value=1)), (2, 2, 0, 4, dict(new_q=0.25, state_mode=2, state_value=1))):
    o.reaction_add_change(rr, side, lvl, old, new, **kw)
ne = o.react()
t = o.get()["type"]
print("events", ne, "angles", o.list_size(al), len(m["angles"]), "quads", o.list_size(ql), np.bincount(t))
b = o.list_get(rl,2)[:5]; print(b, t[b])
