"""Row-block target sweep (home particles per block) on a benchmark workload (run on the GPU box)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from chemlab_b200 import Engine, synthetic

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which == "c2":
    wl = bench.WorkloadC2(int(os.environ.get("NSIDE", "100")))
else:
    wl = synthetic.make_workload(which, int(sys.argv[2]) if len(sys.argv) > 2 else 0, example_root=os.path.join(bench.ROOT, "tests", "golden"))
sysd = wl.system()
e = Engine(sysd["box"], wl.rc, wl.skin, seed=bench.SEED)
bench.upload(e, sysd)
h = wl.setup(e, sysd)
e.reaction_general(0, wl.interval, 1, 0)
e.run(30)
print("## %s  n=%d" % (wl.description, sysd["n"]), flush=True)
V = [("legacy bx8", dict(block_target=-1, block_cells=8)), ("t96", dict(block_target=96, block_cells=16)), ("t128", dict(block_target=128)), ("t160", dict(block_target=160)),
     ("t192", dict(block_target=192)), ("t224", dict(block_target=224)), ("t256", dict(block_target=256)), ("auto", dict(block_target=0))]
for name, opts in V:
    try:
        for k, v in opts.items():
            e.set_option(k, v)
        e.run(6)
        e.reset_timers(); e.set_option("pair_event_timing", 1)
        e.run(60)
        tm, cn = e.timers()
        pm = e.get_option("pair_kernel_ms") / max(1, e.get_option("pair_kernel_launches"))
        e.set_option("pair_event_timing", 0)
        info = {k: e.get_option(k) for k in ("pair_nv", "pair_threads", "pair_smem", "block_cells", "block_target", "blocks", "tile_max", "home_max")}
        print("%-14s pair %.4f ms  step %.4f ms  rebuilds %d  %s" % (name, pm, 1e3 * tm["total"] / 60, cn["rebuilds"], json.dumps(info)), flush=True)
    except Exception as ex:
        print("%-14s FAILED %s" % (name, ex), flush=True)
e.close()
