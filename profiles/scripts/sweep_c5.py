"""Many-table pair kernel: block target x table budget sweep (run on the GPU box)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from chemlab_b200 import Engine, synthetic
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
wl = synthetic.make_workload(which, int(sys.argv[2]) if len(sys.argv) > 2 else 10, example_root=os.path.join(bench.ROOT, "tests", "golden"))
sysd = wl.system()
e = Engine(sysd["box"], wl.rc, wl.skin, seed=bench.SEED)
bench.upload(e, sysd)
h = wl.setup(e, sysd)
e.reaction_general(0, wl.interval, 1, 0)
e.run(30)
print("## %s  n=%d" % (wl.description, sysd["n"]), flush=True)
V = [("auto", dict(block_target=0, pair_table_kb=-1, pair_nv=0))]
for tgt in (96, 128, 160, 192, 224):
    for kb in (-1, 60, 90):
        V.append(("t%d kb%d" % (tgt, kb), dict(block_target=tgt, pair_table_kb=kb, pair_nv=0)))
V += [("legacy bx8", dict(block_target=-1, block_cells=8, pair_table_kb=-1)), ("legacy bx8 kb90", dict(block_target=-1, block_cells=8, pair_table_kb=90))]
for name, opts in V:
    try:
        for k, v in opts.items():
            e.set_option(k, v)
        e.run(6)
        e.reset_timers(); e.set_option("pair_event_timing", 1)
        e.run(40)
        tm, cn = e.timers()
        pm = e.get_option("pair_kernel_ms") / max(1, e.get_option("pair_kernel_launches"))
        e.set_option("pair_event_timing", 0)
        info = {k: e.get_option(k) for k in ("pair_nv", "pair_threads", "pair_smem", "block_target", "blocks", "tile_max", "home_max", "pair_tables_resident", "pair_tables_resident_weight", "pair_table_rows")}
        print("%-18s pair %.4f ms  step %.4f ms  rebuilds %d  %s" % (name, pm, 1e3 * tm["total"] / 40, cn["rebuilds"], json.dumps(info)), flush=True)
    except Exception as ex:
        print("%-18s FAILED %s" % (name, ex), flush=True)
e.close()
