"""espressopp-compatible surface of the B200 engine -- the drop-in boundary of SURVEY 8b.

Only the names chemlab's driver uses are provided (inventory: SURVEY 2.3).  Each class cites the chemlab call site it
serves; arithmetic happens in the CUDA engine behind include/chemlab_b200.h, never here."""
import numpy as np

from ._context import Context, not_in_scope
from . import esutil, bc, storage, interaction, integrator, analysis, io, tools  # noqa: F401,E402

pmi = None


class Version:
    """espressopp.Version().info(): printed by the example scripts (examples/atrp_lj/polymer_melt.py:48)."""
    name = "chemlab_b200.espressopp"

    def info(self):
        return "%s: espressopp surface of the B200 reactive-MD engine (C-ABI include/chemlab_b200.h)" % self.name


class _CommWorld:
    """MPI.COMM_WORLD.size / .rank as the driver and user scripts read them (src/start_simulation.py:155-157,998,1078): one rank
    per GPU under torchrun, else a single rank."""
    @property
    def size(self):
        try:
            import torch.distributed as dist
            return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        except Exception:
            return 1

    @property
    def rank(self):
        try:
            import torch.distributed as dist
            return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        except Exception:
            return 0


class _MPI:
    COMM_WORLD = _CommWorld()


MPI = _MPI()


class Real3D(tuple):
    """espressopp.Real3D(x, y, z): indexable 3-vector (files_io.py:268-279 reads pos[0..2])."""
    def __new__(cls, x=0.0, y=0.0, z=0.0):
        if hasattr(x, "__len__"):
            x, y, z = x
        return tuple.__new__(cls, (float(x), float(y), float(z)))

    x = property(lambda s: s[0]); y = property(lambda s: s[1]); z = property(lambda s: s[2])

    def __add__(s, o): return Real3D(s[0] + o[0], s[1] + o[1], s[2] + o[2])
    def __sub__(s, o): return Real3D(s[0] - o[0], s[1] - o[1], s[2] - o[2])
    def __mul__(s, a): return Real3D(s[0] * a, s[1] * a, s[2] * a)
    __rmul__ = __mul__
    def sqr(s): return s[0] ** 2 + s[1] ** 2 + s[2] ** 2


Int3D = Real3D


class System:
    """espressopp.System(): src/start_simulation.py:148-167,212."""
    def __init__(self):
        self._ctx = Context()
        self._rng = None
        self._bc = None
        self.storage = None
        self.integrator = None
        self.topology_manager = None

    rng = property(lambda s: s._rng)

    @rng.setter
    def rng(self, r):
        self._rng = r
        self._ctx.seed = int(getattr(r, "seed", 0))

    skin = property(lambda s: s._ctx.skin)

    @skin.setter
    def skin(self, v):
        self._ctx.skin = float(v)

    bc = property(lambda s: s._bc)

    @bc.setter
    def bc(self, b):
        self._bc = b
        self._ctx.box = np.asarray(b.boxL, float)

    def addInteraction(self, inter, label=None):
        label = label if label is not None else "interaction_%d" % len(self._ctx.interactions)
        self._ctx.interactions.append((inter, label))
        if self._ctx.engine is not None:
            inter._attach(self._ctx.engine)

    def getAllInteractions(self):
        return {label: inter for inter, label in self._ctx.interactions}

    def getNumberOfInteractions(self):
        return len(self._ctx.interactions)

    def getInteraction(self, k):
        return self._ctx.interactions[k][0]

    def getNameOfInteraction(self, k):
        return self._ctx.interactions[k][1]


class DynamicExcludeList:
    """DynamicExcludeList(integrator, exclusions) + observe_*: src/start_simulation.py:189,378-391,428-441."""
    def __init__(self, integrator_, exclusionlist=None):
        self._ctx = integrator_._system._ctx
        self._ctx.exclusions = [tuple(int(v) for v in p) for p in (exclusionlist or [])]
        if self._ctx.engine is not None:
            self._ctx.engine.set_exclusions(np.asarray(self._ctx.exclusions, np.int64).reshape(-1, 2))

    def _observe(self, lst):
        self._ctx.exclude_observed.append(lst)
        if self._ctx.engine is not None:
            self._ctx.engine.exclusions_observe(lst._h)

    observe_tuple = observe_triple = observe_quadruple = _observe

    def exclude(self, a, b):
        self._ctx.exclusions.append((int(a), int(b)))
        if self._ctx.engine is not None:
            self._ctx.engine.set_exclusions(np.asarray(self.get_list(), np.int64).reshape(-1, 2))

    def get_list(self):
        if self._ctx.engine is not None:
            return [tuple(p) for p in self._ctx.engine.get_exclusions().tolist()]
        return list(self._ctx.exclusions)

    @property
    def size(self):
        return len(self.get_list())


class VerletList:
    """VerletList(system, cutoff=, exclusionlist=): src/start_simulation.py:193-197."""
    def __init__(self, system, cutoff, exclusionlist=None, **kw):
        self._system = system
        self.cutoff = float(cutoff)
        system._ctx.rc = max(system._ctx.rc or 0.0, self.cutoff)
        self.exclusionlist = exclusionlist

    def get_timers(self):
        """list per rank of (name, seconds) pairs, averaged by the driver (src/start_simulation.py:1076, src/tools.py:82-99)"""
        e = self._system._ctx.engine
        if e is None:
            return []
        t, c = e.timers()
        return [[("timeRebuild", float(t.get("neighbour", 0.0))), ("rebuilds", float(c.get("rebuilds", 0)))]]

    def totalSize(self):
        return len(self._system._ctx.require_engine().pairs())

    def getAllPairs(self):
        return [tuple(p) for p in self._system._ctx.require_engine().pairs().tolist()]


class _FixedList:
    arity = 2

    def __init__(self, storage_, *a, **k):
        self._ctx = storage_._system._ctx
        self._h = None
        self._pending = []
        self._ctx.lists.append(self)
        if self._ctx.engine is not None:
            self._attach(self._ctx.engine)

    def _attach(self, e):
        if self._h is None:
            self._h = e.add_list(self.arity)
            if self._pending:
                e.list_add(self._h, np.asarray(self._pending, np.int64).reshape(-1, self.arity))
                self._pending = []

    def _add(self, tuples):
        tuples = [tuple(int(v) for v in t) for t in tuples]
        if self._h is None:
            self._pending.extend(tuples)
        elif tuples:
            self._ctx.engine.list_add(self._h, np.asarray(tuples, np.int64).reshape(-1, self.arity))

    def _all(self):
        if self._h is None:
            return list(self._pending)
        return [tuple(t) for t in self._ctx.engine.list_get(self._h, self.arity).tolist()]

    def totalSize(self):
        return len(self._pending) if self._h is None else self._ctx.engine.list_size(self._h)

    size = totalSize


class FixedPairList(_FixedList):
    """FixedPairList(storage): gromacs_topology.py:1019; reaction_setup.py:449."""
    arity = 2
    addBonds = _FixedList._add
    getAllBonds = getBonds = _FixedList._all

    def add(self, a, b):
        self._add([(a, b)])


class FixedTripleList(_FixedList):
    arity = 3
    addTriples = _FixedList._add
    getAllTriples = getTriples = _FixedList._all

    def add(self, a, b, c):
        self._add([(a, b, c)])


class FixedQuadrupleList(_FixedList):
    arity = 4
    addQuadruples = _FixedList._add
    getAllQuadruples = getQuadruples = _FixedList._all

    def add(self, a, b, c, d):
        self._add([(a, b, c, d)])


FixedPairListLambda = not_in_scope("FixedPairListLambda")
FixedTripleListLambda = not_in_scope("FixedTripleListLambda")
FixedQuadrupleListLambda = not_in_scope("FixedQuadrupleListLambda")
ParticleRegion = not_in_scope("ParticleRegion")
ParticleGroup = not_in_scope("ParticleGroup")
