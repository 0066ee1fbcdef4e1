"""A/B of rebuild / pair options on the C2 melt (run on the GPU box): per-rebuild time from the engine's neighbour bucket."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from chemlab_b200 import Engine, synthetic

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WorkloadC2(100) if which == "c2" else synthetic.make_workload(which, int(sys.argv[2]) if len(sys.argv) > 2 else 0, example_root=os.path.join(bench.ROOT, "tests", "golden"))
sysd = wl.system()
e = Engine(sysd["box"], wl.rc, wl.skin, seed=bench.SEED)
bench.upload(e, sysd)
h = wl.setup(e, sysd)
e.reaction_general(0, wl.interval, 1, 0)
e.set_option("timers", 1)
e.run(30)
print("## %s  n=%d" % (wl.description, sysd["n"]), flush=True)
ref_pairs = None
V = [("build1 perm0", dict(build_kernel=1, pair_perm=0)), ("build2 perm0", dict(build_kernel=2, pair_perm=0)), ("build2 perm1", dict(build_kernel=2, pair_perm=1)),
     ("build1 perm1", dict(build_kernel=1, pair_perm=1)), ("build2 perm1 again", dict(build_kernel=2, pair_perm=1))]
for name, opts in V:
    for k, v in opts.items():
        e.set_option(k, v)
    e.run(6)
    e.reset_timers(); e.set_option("pair_event_timing", 1)
    e.run(120)
    tm, cn = e.timers()
    pm = e.get_option("pair_kernel_ms") / max(1, e.get_option("pair_kernel_launches"))
    e.set_option("pair_event_timing", 0)
    print("%-20s pair %.4f ms  step %.4f ms  rebuilds %d  neighbour bucket %.3f ms/rebuild  bonded %.4f integ %.4f ms/step  build_kernel %d" %
          (name, pm, 1e3 * tm["total"] / 120, cn["rebuilds"], 1e3 * tm["neighbour"] / max(1, cn["rebuilds"]), 1e3 * tm["bonded"] / 120, 1e3 * tm["integrate"] / 120,
           e.get_option("build_kernel")), flush=True)
# pair sets of the two build kernels on the same positions
e.set_option("build_kernel", 1); p1 = e.pairs()
e.set_option("build_kernel", 2); p2 = e.pairs()
print("pair sets equal:", p1.shape == p2.shape and bool((p1 == p2).all()), len(p1), flush=True)
e.close()
