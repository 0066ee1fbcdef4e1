"""tools.analyse.final_info: src/start_simulation.py:1078-1079; tools.analyse.info: the one-line status print of the example
scripts (examples/atrp_lj/polymer_melt.py:126-157)."""


def final_info(system, integrator, vl=None, start_time=None, end_time=None):
    e = system._ctx.engine
    if e is None:
        return
    t, c = e.timers()
    print("steps=%d rebuilds=%d launches=%d" % (c.get("steps", 0), c.get("rebuilds", 0), c.get("launches", 0)))
    for k in ("pair", "bonded", "neighbour", "integrate", "comm", "reaction", "total"):
        print("  %-10s %10.4f s" % (k, t.get(k, 0.0)))


def info(system, integrator, per_atom=False):
    """step, temperature, kinetic energy and the energy of every registered interaction, one line."""
    e = system._ctx.require_engine()
    ek, T, n = (float(x) for x in e.kinetics())
    scale = 1.0 / n if (per_atom and n) else 1.0
    cols = ["step=%d" % integrator.step, "T=%.6g" % T, "Ekin=%.6g" % (ek * scale)]
    tot = ek
    for k in range(system.getNumberOfInteractions()):
        inter = system.getInteraction(k)
        if getattr(inter, "_zero_energy", False):
            continue
        if getattr(inter, "_h", None) is None:
            inter._attach(e)
        v = float(e.energy(inter._h))
        tot += v
        cols.append("%s=%.6g" % (system.getNameOfInteraction(k), v * scale))
    cols.append("Etotal=%.6g" % (tot * scale))
    print(" ".join(cols))
