"""espressopp.io: DumpGRO is provided; DumpH5MD / DumpTopology need h5py (absent here) and are out of scope."""
from ._context import not_in_scope


class DumpGRO:
    """io.DumpGRO(system, integrator, filename=, unfolded=, append=): src/start_simulation.py:684-696."""
    def __init__(self, system, integrator_, filename="out.gro", unfolded=False, append=True, **kw):
        self._system, self._integrator = system, integrator_
        self.filename, self.unfolded, self.append = filename, unfolded, append
        self._first = True

    def dump(self):
        ctx = self._system._ctx
        g = ctx.require_engine().get_particles(fields=("pos", "image"))
        pos = g["pos"] + (g["image"] * ctx.box if self.unfolded else 0.0)
        mode = "a" if (self.append and not self._first) else "w"
        self._first = False
        with open(self.filename, mode) as f:
            f.write("chemlab_b200 step %d\n%d\n" % (self._integrator.step, len(pos)))
            for k, pid in enumerate(sorted(ctx.pid)):
                f.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (1, "MOL", "A", pid % 100000, pos[k, 0], pos[k, 1], pos[k, 2]))
            f.write("%10.5f%10.5f%10.5f\n" % tuple(ctx.box))

    perform_action = dump


DumpH5MD = not_in_scope("io.DumpH5MD")
DumpTopology = not_in_scope("io.DumpTopology")
DumpXYZ = not_in_scope("io.DumpXYZ")
