"""storage.DomainDecomposition(system, nodeGrid, cellGrid): src/start_simulation.py:163-171.
The MPI node grid becomes the engine's slab decomposition (clb_comm_init); the cell grid is derived by the engine."""
import numpy as np


class Particle:
    """Result of storage.getParticle(pid): .pos .v .f .type .mass .q .state .res_id .imageBox
    (src/start_simulation.py:855-873; files_io.py:268-279)."""
    def __init__(self, pid, d):
        from . import Real3D
        self.id = pid
        self.pos = Real3D(*d["pos"])
        self.v = Real3D(*d["vel"])
        self.f = Real3D(*d.get("force", (0, 0, 0)))
        self.type = int(d["type"]); self.mass = float(d["mass"]); self.q = float(d["q"])
        self.state = int(d["state"]); self.res_id = int(d["res_id"])
        self.imageBox = tuple(int(x) for x in d["image"])
        self.lambda_adr = 0.0


class DomainDecomposition:
    def __init__(self, system, nodeGrid=None, cellGrid=None, *a, **k):
        self._system = system
        self.nodeGrid, self.cellGrid = nodeGrid, cellGrid

    def addParticles(self, plist, *props):
        self._system._ctx.add_particles(plist, props)

    def addParticle(self, pid, pos):
        self._system._ctx.add_particles([(pid, tuple(pos))], ("id", "pos"))

    def decompose(self):
        ctx = self._system._ctx
        if ctx.engine is not None:
            ctx.engine.decompose()

    def particleExists(self, pid):
        return int(pid) in self._system._ctx.pid_index

    def getParticle(self, pid):
        ctx = self._system._ctx
        pid = int(pid)
        if pid not in ctx.pid_index:
            raise RuntimeError("particle %d does not exist" % pid)
        if ctx.engine is None:
            k = ctx.pid_index[pid]; P = ctx.props
            d = dict(pos=P["pos"][k], vel=P["v"][k], type=P["type"][k], mass=P["mass"][k], q=P["q"][k], state=P["state"][k],
                     res_id=P["res_id"][k], image=(0, 0, 0))
        else:
            g = ctx.engine.get_particles(ids=[pid])
            d = {k: (v[0] if hasattr(v, "__len__") else v) for k, v in g.items()}
        return Particle(pid, d)

    _FIELD = {"type": "type", "state": "state", "mass": "mass", "q": "q", "res_id": "res_id", "pos": "pos", "v": "v"}

    def modifyParticle(self, pid, prop, value):
        ctx = self._system._ctx
        pid = int(pid)
        if prop not in self._FIELD:
            raise RuntimeError("modifyParticle: property %s is not supported" % prop)
        if ctx.engine is None:
            k = ctx.pid_index[pid]
            ctx.props[prop][k] = tuple(value) if prop in ("pos", "v") else value
        else:
            ctx.engine.modify_particle(pid, prop, np.atleast_1d(np.asarray(value, float)))

    def getAllParticleIDs(self):
        return list(self._system._ctx.pid)
