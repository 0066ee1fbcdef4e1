"""espressopp.analysis.* observables chemlab registers with its SystemMonitor (src/start_simulation.py:446-569)."""
import numpy as np

from ._context import not_in_scope


class _Obs:
    def __init__(self, system, *a, **k):
        self._system = system
        self._ctx = system._ctx


class Temperature(_Obs):
    """analysis.Temperature(system).compute(): T = 2 Ekin / (3 N) in energy units (kB folded into the thermostat value)."""
    def add_type(self, t):
        pass

    def compute(self):
        return float(self._ctx.require_engine().kinetics()[1])


class KineticEnergy(_Obs):
    def __init__(self, system, temperature=None):
        super().__init__(system)

    def compute(self):
        return float(self._ctx.require_engine().kinetics()[0])


class PotentialEnergy(_Obs):
    """analysis.PotentialEnergy(system, interaction): src/start_simulation.py:470-480."""
    def __init__(self, system, interaction):
        super().__init__(system)
        self._inter = interaction

    def compute(self):
        e = self._ctx.require_engine()
        if getattr(self._inter, "_zero_energy", False):         # the `coulomb` term of an all-neutral system
            return 0.0
        if getattr(self._inter, "_h", None) is None:
            self._inter._attach(e)
        return float(e.energy(self._inter._h))


class NPart(_Obs):
    def compute(self):
        return len(self._ctx.pid)


class MaxPID(_Obs):
    def compute(self):
        return max(self._ctx.pid)


class _NEntries(_Obs):
    def __init__(self, system, flist):
        super().__init__(system)
        self._list = flist

    def compute(self):
        return int(self._list.totalSize())


NFixedPairListEntries = NFixedTripleListEntries = NFixedQuadrupleListEntries = _NEntries


class NExcludeListEntries(_Obs):
    def __init__(self, system, vl):
        super().__init__(system)

    def compute(self):
        e = self._ctx.engine
        return len(self._ctx.exclusions) if e is None else len(e.get_exclusions())


class ChemicalConversion(_Obs):
    """ChemicalConversion(system, type, total): N(type)/total (gromacs_topology.py:574-583; tools.py:102-180)."""
    def __init__(self, system, type_id, total=None):
        super().__init__(system)
        self.type_id, self.total = int(type_id), (float(total) if total else 1.0)

    def compute(self):
        return self._ctx.require_engine().count_type(self.type_id) / self.total


class ChemicalConversionTypeState(_Obs):
    """ChemicalConversionTypeState(system, type, state, total): N(type in state)/total; or, as src/tools.py:142-158 uses it for
    'A+B(1)+...' stop criteria, ChemicalConversionTypeState(system, total_count=total) followed by count_type(type, state | None)
    for every summand."""
    def __init__(self, system, type_id=None, state=None, total=None, total_count=None):
        super().__init__(system)
        tot = total if total is not None else total_count
        self.total = float(tot) if tot else 1.0
        self._counted = []
        if type_id is not None:
            self.count_type(type_id, state)

    def count_type(self, type_id, state=None):
        self._counted.append((int(type_id), -1 if state is None else int(state)))

    def compute(self):
        e = self._ctx.require_engine()
        return sum(e.count_type(t, s) for t, s in self._counted) / self.total


class AngleDistribution(_Obs):
    """analysis.AngleDistribution(system) + .load_from_topology_manager(tm) + .compute(nbins): the user hook of
    examples/pccg_lj/chemical_reactions/hooks.py:44-56 accumulates it at every outer step.  [EXT] semantics restated (U28): every
    angle i-j-k spanned by two bonds that meet at j in the TopologyManager's bond graph (the observed pair lists, reaction bonds
    included) is binned over [0, pi) into `nbins` equal bins; the raw counts are returned.  An analysis observable outside the hot
    path: bonds and positions are read back through the C-ABI (clb_list_get, clb_get_particles) and binned with numpy."""
    def __init__(self, system):
        super().__init__(system)
        self._tm = None

    def load_from_topology_manager(self, tm):
        self._tm = tm

    def compute(self, nbins):
        nbins = int(nbins)
        hist = np.zeros(nbins, np.int64)
        tm = self._tm if self._tm is not None else getattr(self._ctx, "topology_manager", None)
        if tm is None:
            return hist.tolist()
        e = self._ctx.require_engine()
        bonds = [np.asarray(f.getAllBonds(), np.int64).reshape(-1, 2) for f in tm._observed]
        bonds = np.concatenate(bonds) if bonds else np.zeros((0, 2), np.int64)
        if len(bonds) == 0:
            return hist.tolist()
        ids, inv = np.unique(bonds, return_inverse=True)
        inv = inv.reshape(-1, 2)
        pos = np.asarray(e.get_particles(ids=ids, fields=("pos",))["pos"], float)
        box = np.asarray(self._system.bc.boxL, float)
        # half-edges sorted by their centre j; all unordered pairs of half-edges of one centre span an angle
        centre = np.concatenate([inv[:, 0], inv[:, 1]]); other = np.concatenate([inv[:, 1], inv[:, 0]])
        order = np.argsort(centre, kind="stable")
        centre, other = centre[order], other[order]
        start = np.flatnonzero(np.r_[True, centre[1:] != centre[:-1]])
        count = np.diff(np.r_[start, len(centre)])
        for deg in np.unique(count[count >= 2]):
            rows = start[count == deg]
            nb = other[rows[:, None] + np.arange(deg)[None, :]]                 # [centres][deg]
            c = pos[centre[rows]]
            v = pos[nb] - c[:, None, :]
            v -= box * np.rint(v / box)
            v /= np.linalg.norm(v, axis=2, keepdims=True)
            a, b = np.triu_indices(deg, 1)
            cos = np.clip(np.einsum("nkd,nkd->nk", v[:, a, :], v[:, b, :]), -1.0, 1.0)
            hist += np.histogram(np.arccos(cos).ravel(), bins=nbins, range=(0.0, np.pi))[0]
        return hist.tolist()


class CMVelocity(_Obs):
    """analysis.CMVelocity(system).reset(): removes the centre-of-mass velocity (src/start_simulation.py:680-682)."""
    def compute(self):
        e = self._ctx.require_engine()
        g = e.get_particles(fields=("vel", "mass"))
        return tuple((g["vel"] * g["mass"][:, None]).sum(0) / g["mass"].sum())

    def reset(self):
        e = self._ctx.require_engine()
        g = e.get_particles(fields=("vel", "mass"))
        v = g["vel"] - (g["vel"] * g["mass"][:, None]).sum(0) / g["mass"].sum()
        e.set_velocities(v)


class SystemMonitorOutputCSV:
    def __init__(self, filename, delimiter="\t"):
        self.filename, self.delimiter = filename, delimiter
        self._fh = None

    def write(self, header, row):
        import os
        if int(os.environ.get("RANK", "0")) != 0:      # multi-GPU: every rank computes (collectives), rank 0 writes
            return
        if self._fh is None:
            self._fh = open(self.filename, "w")
            self._fh.write(self.delimiter.join(header) + "\n")
        self._fh.write(self.delimiter.join("%.10g" % v if isinstance(v, float) else str(v) for v in row) + "\n")
        self._fh.flush()


class SystemMonitor:
    """SystemMonitor(system, integrator, output) + add_observable / info / dump: src/start_simulation.py:446-569,729."""
    def __init__(self, system, integrator_, output):
        self._system, self._integrator, self._out = system, integrator_, output
        self._obs = []
        self.potential_energy = 0.0
        self._last = None

    def add_observable(self, name, obs, visible=True):
        self._obs.append((name, obs, visible))

    def perform_action(self):
        step = self._integrator.step
        vals = [obs.compute() for _, obs, _ in self._obs]
        self.potential_energy = float(sum(v for (n, o, _), v in zip(self._obs, vals) if isinstance(o, PotentialEnergy)))
        self._last = (step, vals)
        self._out.write(["step", "time"] + [n for n, _, _ in self._obs], [step, step * self._integrator.dt] + vals)

    dump = perform_action

    def info(self):
        if self._last is None:
            return
        step, vals = self._last
        cols = ["%s=%.6g" % (n, v) for (n, _, vis), v in zip(self._obs, vals) if vis]
        print("step %d: %s" % (step, " ".join(cols)))


for _n in ("Pressure", "BoxSize", "ResolutionFixedPairList", "NParticlePairScalingEntries", "NumFixDistances", "MaxForce"):
    globals()[_n] = not_in_scope("analysis." + _n)
del _n
