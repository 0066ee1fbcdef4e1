"""Shared state behind the espressopp-style objects: host-side declarations until the first operation that
needs the GPU, then one chemlab_b200.Engine (C-ABI) that every object talks to through its handle."""
import numpy as np

from ..engine import Engine, EngineError  # noqa: F401


class Context:
    def __init__(self):
        self.engine = None
        self.box = None
        self.skin = 0.0
        self.seed = 0
        self.rc = None
        # particles as given by storage.addParticles (host arrays until the engine exists)
        self.pid = []
        self.props = {}                 # name -> list
        self.pid_index = {}
        self.exclusions = []
        self.exclude_observed = []      # tuple lists observed by the DynamicExcludeList
        self.lists = []                 # FixedXList objects in creation order
        self.interactions = []          # (interaction object, label) in system.addInteraction order
        self.tables = {}                # filename/itype -> engine handle
        self.integrator = None
        self.topology_manager = None
        self.dirty = True

    # ---- particles on the host
    def add_particles(self, rows, names):
        for r in rows:
            d = dict(zip(names, r))
            pid = int(d["id"])
            if pid in self.pid_index:
                raise RuntimeError("particle id %d added twice" % pid)
            self.pid_index[pid] = len(self.pid)
            self.pid.append(pid)
            for k in ("type", "pos", "v", "mass", "q", "state", "res_id"):
                default = {"pos": (0.0, 0.0, 0.0), "v": (0.0, 0.0, 0.0), "mass": 1.0, "q": 0.0}.get(k, 0)
                v = d.get(k, default)
                if k in ("pos", "v"):
                    v = tuple(float(x) for x in v)
                self.props.setdefault(k, []).append(v)
        self.dirty = True

    def require_engine(self):
        if self.engine is None:
            self.build()
        return self.engine

    def build(self):
        if self.box is None or self.rc is None:
            raise RuntimeError("System needs system.bc (box) and a VerletList (cutoff) before it can run")
        if not self.pid:
            raise RuntimeError("no particles: call storage.addParticles first")
        # under torchrun (the mpirun of this engine) every rank builds the same system and joins the slab decomposition
        device, join = 0, False
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                import os
                device, join = int(os.environ.get("LOCAL_RANK", "0")), True
        except ImportError:
            pass
        e = Engine(self.box, self.rc, self.skin, seed=self.seed, device=device)
        if join:
            e.join()
        P = self.props
        e.set_particles(np.asarray(self.pid, np.int64), np.asarray(P["type"], np.int32), np.asarray(P["pos"], float),
                        np.asarray(P["mass"], float), vel=np.asarray(P["v"], float), q=np.asarray(P["q"], float),
                        state=np.asarray(P["state"], np.int32), res_id=np.asarray(P["res_id"], np.int32))
        self.engine = e
        if self.exclusions:
            e.set_exclusions(np.asarray(self.exclusions, np.int64).reshape(-1, 2))
        for lst in self.lists:
            lst._attach(e)
        for inter, label in self.interactions:
            inter._attach(e)
        for lst in self.exclude_observed:
            e.exclusions_observe(lst._h)
        if self.topology_manager is not None:
            self.topology_manager._attach(e)
        if self.integrator is not None:
            self.integrator._attach(e)
        self.dirty = False

    def table(self, pot):
        """engine table handle of a Tabulated* potential (one upload per file and interpolation type)"""
        key = (pot.filename, pot.itype)
        if key not in self.tables:
            from .interaction import read_pot
            r, en, f = read_pot(pot.filename)
            self.tables[key] = self.engine.add_table(r, en, f, pot.itype)
        return self.tables[key]


def not_in_scope(name):
    """Names chemlab constructs unconditionally but this engine does not implement: constructible,
    NotImplementedError on first real use."""
    class _Stub:
        _clb_stub = name

        def __init__(self, *a, **k):
            self._args = (a, k)

        def _attach(self, engine):
            raise NotImplementedError("espressopp.%s is outside the scope of the B200 engine (SURVEY 2.3 E20/E21)" % name)

        def __getattr__(self, item):
            if item.startswith("_"):
                raise AttributeError(item)

            def _f(*a, **k):
                raise NotImplementedError("espressopp.%s.%s is outside the scope of the B200 engine" % (name, item))
            return _f
    _Stub.__name__ = name.split(".")[-1]
    return _Stub
