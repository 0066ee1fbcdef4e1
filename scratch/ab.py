import sys, time, itertools; sys.path.insert(0,'/root/repo')
import numpy as np
from chemlab_b200 import Engine, synthetic
import bench
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
sysd = synthetic.trimer_melt(n_side, rho=bench.RHO, seed=12345, kT=bench.KT)
e, h = bench.build_engine(sysd)
e.run(300)
def timeit(tag, steps=200):
    e.run(20)
    e.reset_timers(); e.set_option("pair_event_timing", 1)
    e.run(steps)
    tm, cn = e.timers()
    pm = e.get_option("pair_kernel_ms") / max(1, e.get_option("pair_kernel_launches"))
    print("%-40s steps/s=%8.1f ms/step=%.3f pair_ms=%.4f rebuilds=%d split=%d threads=%d grid=%d smem=%d" % (tag, steps / tm["total"], 1e3 * tm["total"] / steps, pm, cn["rebuilds"], e.get_option("pair_split"), e.get_option("pair_threads"), e.get_option("pair_grid"), e.get_option("pair_smem")), flush=True)
    e.set_option("pair_event_timing", 0)
for bf, sp in itertools.product((1, 0), (1, 2, 4)):
    e.set_option("pair_branchfree", bf); e.set_option("pair_split", sp)
    timeit("branchfree=%d split=%d" % (bf, sp))
for bx in (4, 6, 12):
    e.set_option("pair_branchfree", 1); e.set_option("pair_split", 0); e.set_option("block_cells", bx)
    timeit("bx=%d auto split" % bx)
