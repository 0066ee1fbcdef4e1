"""pair-kernel pipeline / block target / perm sweep (run on the GPU box)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from chemlab_b200 import Engine, synthetic

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WorkloadC2(100) if which == "c2" else synthetic.make_workload(which, int(sys.argv[2]) if len(sys.argv) > 2 else 0, example_root=os.path.join(bench.ROOT, "tests", "golden"))
sysd = wl.system()
e = Engine(sysd["box"], wl.rc, wl.skin, seed=bench.SEED)
bench.upload(e, sysd)
h = wl.setup(e, sysd)
e.reaction_general(0, wl.interval, 1, 0)
e.run(30)
print("## %s  n=%d" % (wl.description, sysd["n"]), flush=True)
V = []
for pipe in (0, 1):
    for tgt in (96, 128, 160, 192):
        for perm in (1, 0):
            V.append(("pipe%d t%d perm%d" % (pipe, tgt, perm), dict(pair_pipe=pipe, block_target=tgt, pair_perm=perm, pair_nv=0)))
V += [("pipe1 t128 nv3", dict(pair_pipe=1, block_target=128, pair_perm=1, pair_nv=3)), ("pipe1 t96 nv4", dict(pair_pipe=1, block_target=96, pair_perm=1, pair_nv=4)),
      ("pipe1 t64 perm1", dict(pair_pipe=1, block_target=64, pair_perm=1, pair_nv=0)), ("auto", dict(pair_pipe=-1, block_target=0, pair_perm=1, pair_nv=0))]
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
        info = {k: e.get_option(k) for k in ("pair_nv", "pair_threads", "pair_smem", "block_target", "blocks", "tile_max", "home_max", "pair_pipe")}
        print("%-22s pair %.4f ms  step %.4f ms  rebuilds %d  %s" % (name, pm, 1e3 * tm["total"] / 60, cn["rebuilds"], json.dumps(info)), flush=True)
    except Exception as ex:
        print("%-22s FAILED %s" % (name, ex), flush=True)
e.close()
