import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, clb_testutil as util
m = util.melt(12, seed=1); n=len(m['pos'])
P = util.Pair(m['pos'], m['box'], m['type'])
r,e,f = util.lj_table(); tab = P.add_table(r,e,f,1); nb = P.nb_tab(util.type_pairs(2), tab, 2.5)
P.e.set_option("pair_split", 1)
P.e.compute_forces(); P.o.compute_forces()
fe = P.e.get_particles(fields=("force",))["force"]; fo = P.o.get()["force"]
d = np.linalg.norm(fe-fo,axis=1); bad = d > 1e-6*np.linalg.norm(fo,axis=1).mean()
st = P.e.get_particles(fields=("pos",))["pos"]; L = m['box'][0]
print("nbad", bad.sum(), "of", n)
frac = st/L
for dim in range(3):
    print("dim", dim, "bad frac range", frac[bad][:,dim].min(), frac[bad][:,dim].max(), "good range", frac[~bad][:,dim].min(), frac[~bad][:,dim].max())
print("zero forces among bad:", (np.abs(fe[bad]).max(axis=1)==0).sum())
nz = bad & (np.abs(fe).max(axis=1)>0)
print("nonzero-but-wrong", nz.sum(), fe[nz][:3], fo[nz][:3])
