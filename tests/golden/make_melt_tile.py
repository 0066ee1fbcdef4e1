#!/usr/bin/env python
"""Generates chemlab_b200/data/melt_tile_20.npz: an EQUILIBRATED 8000-bead tile of the C2 trimer melt
(BASELINE.json configs[1]) that `chemlab_b200.synthetic.replicated_melt` replicates
periodically to the benchmark size (5x5x5 tiles = 1,000,000 beads).  SURVEY 8d: "directly a pre-equilibrated
replicated tile of a small box".  Both bench arms (GPU engine and the CPU reference arm) start from this
same state, so neither needs an untimed equilibration run.

Run once, by hand:  python tests/golden/make_melt_tile.py   (about a minute on 8 cores; uses the CPU oracle,
which is why this script lives under tests/ and not in the product package)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from chemlab_b200 import synthetic  # noqa: E402
from oracle import pyoracle  # noqa: E402

N_SIDE, STEPS = 20, 30000


def main():
    s = synthetic.trimer_melt(N_SIDE, rho=0.8442, seed=4242, kT=1.0, vel_seed=4243)
    o = pyoracle.Oracle(s["n"], s["box"], 2.5, 0.3, seed=4244)
    o.set_threads(o.max_threads())
    o.set_particles(s["pos"], s["vel"], s["mass"], None, s["type"], s["state"], s["resid"])
    synthetic.setup_reactive_melt(o, s, rc=2.5, dt=0.005, kT=1.0, gamma=1.0, reactions=False)
    done = 0
    while done < STEPS:
        o.run(5000); done += 5000
        st = o.get()
        ek = 0.5 * (s["mass"][:, None] * st["vel"] ** 2).sum()
        print("step %d  T = %.4f" % (done, 2 * ek / (3 * s["n"])), flush=True)
    st = o.get()
    pos = st["pos"] + st["image"] * s["box"]           # unfolded: bonded beads stay next to each other
    out = os.path.join(ROOT, "chemlab_b200", "data", "melt_tile_20.npz")
    np.savez_compressed(out, n_side=N_SIDE, box=s["box"], pos=pos, vel=st["vel"], type=s["type"], state=s["state"],
                        resid=s["resid"], bonds=s["bonds"], angles=s["angles"], exclusions=s["exclusions"], steps=STEPS)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
