"""tools.analyse.final_info: src/start_simulation.py:1078-1079."""


def final_info(system, integrator, vl=None, start_time=None, end_time=None):
    e = system._ctx.engine
    if e is None:
        return
    t, c = e.timers()
    print("steps=%d rebuilds=%d launches=%d" % (c["steps"], c["rebuilds"], c["launches"]))
    for k in ("pair", "bonded", "neighbour", "integrate", "comm", "reaction", "total"):
        print("  %-10s %10.4f s" % (k, t[k]))
