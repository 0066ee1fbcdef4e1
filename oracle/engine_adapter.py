"""Engine-shaped adapter over the fp64 CPU oracle (TEST INFRASTRUCTURE: tests/, smoke() and bench.py's CPU legs only):
lets the espressopp surface, the start_simulation driver and the recorded workloads of bench.py run unchanged on the
oracle, so that a whole chemlab run on the GPU engine can be compared with the same run on the checker.  Maps caller
ids <-> the oracle's dense indices (ascending id order)."""
import numpy as np

from oracle import pyoracle
from chemlab_b200.engine import NB, POT


class OracleEngine:
    FIELDS = dict(type=0, state=1, mass=2, q=3, res_id=4, pos=5, v=6, vel=6)

    def __init__(self, box, rc_max, skin, seed=0, device=0):
        self.box = np.asarray(box, float)
        self._args = (rc_max, skin, seed)
        self.o = None

    # ---- particles
    def set_particles(self, ids, type, pos, mass, vel=None, q=None, state=None, res_id=None):
        ids = np.asarray(ids, np.int64)
        order = np.argsort(ids, kind="stable")
        self.ids = ids[order]
        self.n = len(ids)
        # dense ascending ids (the usual .gro numbering): index = id - first id, no dictionary
        self.base = int(self.ids[0]) if self.n and np.array_equal(self.ids, self.ids[0] + np.arange(self.n, dtype=np.int64)) else None
        self.idx = None if self.base is not None else {int(p): k for k, p in enumerate(self.ids)}
        rc, skin, seed = self._args
        self.o = pyoracle.Oracle(self.n, self.box, rc, skin, seed=seed)
        z = np.zeros((self.n, 3))
        self.o.set_particles(np.asarray(pos, float)[order], (np.asarray(vel, float)[order] if vel is not None else z),
                             np.asarray(mass, float)[order], (np.asarray(q, float)[order] if q is not None else None),
                             np.asarray(type, np.int32)[order], (np.asarray(state, np.int32)[order] if state is not None else None),
                             (np.asarray(res_id, np.int32)[order] if res_id is not None else None))
        self._resid = np.asarray(res_id, np.int32)[order] if res_id is not None else np.zeros(self.n, np.int32)
        self._q = np.asarray(q, float)[order] if q is not None else np.zeros(self.n)

    def _ix(self, a):
        a = np.asarray(a, np.int64)
        if self.base is not None:
            if a.size and (a.min() < self.base or a.max() >= self.base + self.n):
                raise KeyError("unknown particle id")
            return a - self.base
        return np.vectorize(self.idx.__getitem__, otypes=[np.int64])(a) if a.size else a

    def num_particles(self):
        return self.n

    def get_particles(self, ids=None, fields=("pos", "vel", "force", "type", "state", "mass", "image", "q", "res_id"), out=None):
        g = self.o.get()
        sel = slice(None) if ids is None else self._ix(ids)
        # the oracle folds at rebuilds only: between two rebuilds a stored coordinate may sit just outside [0, L).  The engine
        # always hands out the folded position with the matching image counter -- do the same here
        shift = np.floor(g["pos"] / self.box)
        folded = g["pos"] - shift * self.box
        wrap = folded >= self.box                                    # x = -1e-17 -> x + L == L in floating point
        folded[wrap] -= np.broadcast_to(self.box, folded.shape)[wrap]
        shift[wrap] += 1
        out = {}
        for f in fields:
            if f == "q":
                out[f] = self._q[sel]
            elif f == "res_id":
                out[f] = self._resid[sel]
            elif f == "pos":
                out[f] = folded[sel]
            elif f == "image":
                out[f] = (g["image"] + shift.astype(g["image"].dtype))[sel]
            else:
                out[f] = g[f][sel]
        return out

    def modify_particle(self, pid, field, value):
        f = self.FIELDS[field] if isinstance(field, str) else int(field)
        k = int(self._ix(np.array([pid]))[0])
        if f == 3:
            self._q[k] = float(np.atleast_1d(value)[0])
        if f == 4:
            self._resid[k] = int(np.atleast_1d(value)[0])
        self.o.modify(k, f, value)

    def set_velocities(self, vel):
        self.o.set_velocities(vel)

    def set_positions(self, pos):
        self.o.set_positions(pos)

    # ---- exclusions / lists carry ids
    def set_exclusions(self, pairs):
        self.o.set_exclusions(self._ix(np.asarray(pairs, np.int64).reshape(-1, 2)))

    def get_exclusions(self):
        return self.ids[self.o.get_exclusions()]

    def reaction_define_connections(self, r, pairs):
        self.o.reaction_define_connections(r, self._ix(np.asarray(pairs, np.int64).reshape(-1, 2)))

    def list_add(self, lst, ids):
        ids = np.asarray(ids, np.int64)
        if ids.size:
            self.o.list_add(lst, self._ix(ids))

    def list_get(self, lst, arity):
        return self.ids[self.o.list_get(lst, arity)]

    def pairs(self):
        return self.ids[self.o.pairs()]

    def pairs_raw(self):
        return self.ids[self.o.pairs_raw()]

    def get(self):
        return self.o.get()

    def step(self):
        return self.o.step()

    def set_threads(self, nt):
        self.o.set_threads(nt)

    def last_candidates(self):
        rows, d2 = self.o.candidates()
        rows = rows.astype(np.int64)
        rows[:, 0] = self.ids[rows[:, 0]]; rows[:, 1] = self.ids[rows[:, 1]]
        return rows, d2

    def add_nonbonded(self, kind):
        return self.o.add_nonbonded(NB[kind] if isinstance(kind, str) else int(kind))

    def bonded_set_potential(self, inter, types, kind, params=(), table=-1):
        self.o.bonded_set_potential(inter, types, POT[kind] if isinstance(kind, str) else int(kind), params, table)

    def reaction_counters(self, n):
        return np.array([self.o.reaction_counter(k) for k in range(n)], np.int64)

    def timers(self):
        return {}, {"steps": self.o.step(), "rebuilds": self.o.nrebuild()}

    def decompose(self):
        self.o.rebuild()

    def set_option(self, name, value):
        self.o.set_option(name, value)       # the oracle knows "resort_criterion" and "step"; everything else is ignored there

    def join(self):
        """multi-rank driver runs on the checker: every rank simulates the whole system (no decomposition on the CPU side)"""

    def close(self):
        pass

    def __getattr__(self, name):      # everything else has the same name and arguments on both sides
        return getattr(self.o, name)
