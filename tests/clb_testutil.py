"""Shared builders for the parity tests: small seeded systems fed identically to oracle and engine."""
import numpy as np


def lj_table(rmax=3.0, dr=0.002, eps=1.0, sig=1.0, rc=2.5):
    """LJ 12-6 tabulated on the grid of examples/atrp_activator/table_MA_MA.pot (dr=0.002 from r=dr)."""
    r = dr * np.arange(1, int(round(rmax / dr)) + 1)
    sr6 = (sig / r) ** 6
    src6 = (sig / rc) ** 6
    e = 4 * eps * (sr6 * sr6 - sr6) - 4 * eps * (src6 * src6 - src6)
    f = 24 * eps * (2 * sr6 * sr6 - sr6) / r
    return r, e, f


def melt(n_side, rho=0.8442, seed=1, jitter=0.08, trimers=True):
    """n_side^3 beads on a jittered simple-cubic lattice, grouped as A-L-A trimers along x.

    Returns dict(pos, box, type, resid, bonds, angles). Types: 0 = A (reactive end), 1 = L (middle)."""
    rng = np.random.default_rng(seed)
    n = n_side ** 3
    L = (n / rho) ** (1.0 / 3.0)
    a = L / n_side
    g = np.arange(n_side)
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.float64) * a + 0.5 * a
    pos += rng.uniform(-jitter, jitter, pos.shape) * a
    type_ = np.zeros(n, np.int32)
    resid = np.arange(n, dtype=np.int32)
    bonds, angles = [], []
    if trimers:
        ntri = n_side // 3
        idx = np.arange(n).reshape(n_side, n_side, n_side)
        rid = 0
        for k in range(n_side):
            for j in range(n_side):
                for t in range(ntri):
                    a0, a1, a2 = idx[k, j, 3 * t], idx[k, j, 3 * t + 1], idx[k, j, 3 * t + 2]
                    bonds += [(a0, a1), (a1, a2)]
                    angles.append((a0, a1, a2))
                    type_[a1] = 1
                    resid[[a0, a1, a2]] = rid
                    rid += 1
                for r in range(3 * ntri, n_side):
                    resid[idx[k, j, r]] = rid
                    rid += 1
    return dict(pos=pos, box=np.array([L, L, L]), type=type_, resid=resid,
                bonds=np.array(bonds, np.int64).reshape(-1, 2), angles=np.array(angles, np.int64).reshape(-1, 3))


def exclusions_from(bonds, angles):
    ex = [tuple(b) for b in bonds] + [(a[0], a[2]) for a in angles]
    return np.array(sorted(set((min(a, b), max(a, b)) for a, b in ex)), np.int64).reshape(-1, 2)


def rel_force_err(f, fref):
    """max_i |f_i - fref_i| / max(|fref_i|, rms|fref|): the norm used for the 1e-6 force bound."""
    d = np.linalg.norm(f - fref, axis=1)
    nr = np.linalg.norm(fref, axis=1)
    rms = np.sqrt((nr ** 2).mean())
    return float((d / np.maximum(nr, rms)).max())


class Pair:
    """The same system set up on the CUDA engine (through the C-ABI) and on the fp64 oracle.

    Positions are pushed to the engine first and read back (the engine stores them on its 2^32
    lattice); the oracle is fed exactly those values, so both sides see identical inputs."""

    def __init__(self, pos, box, type, mass=None, vel=None, state=None, resid=None, rc=2.5, skin=0.3, seed=11, ids=None, join=False, device=0):
        from chemlab_b200 import Engine
        from oracle import pyoracle
        n = len(pos)
        self.n = n
        self.ids = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, np.int64)
        mass = np.ones(n) if mass is None else np.asarray(mass, float)
        state = np.zeros(n, np.int32) if state is None else np.asarray(state, np.int32)
        resid = np.arange(n, dtype=np.int32) if resid is None else np.asarray(resid, np.int32)
        # oracle index k <-> k-th smallest id: feed both sides in ascending-id order
        order = np.argsort(self.ids, kind="stable")
        self.ids = self.ids[order]
        pos = np.asarray(pos, float)[order]; type = np.asarray(type, np.int32)[order]; mass = mass[order]
        state = state[order]; resid = resid[order]
        vel = None if vel is None else np.asarray(vel, float)[order]
        self.e = Engine(box, rc, skin, seed=seed, device=device)
        if join:
            self.e.join()       # multi-GPU: this engine becomes one slab (torch.distributed must be initialised)
        self.e.set_particles(self.ids, type, pos, mass, vel=vel, state=state, res_id=resid)
        st = self.e.get_particles(fields=("pos", "vel", "mass", "image"))
        self.o = pyoracle.Oracle(n, box, rc, skin, seed=seed)
        # unfolded position = folded + image * L so that the oracle reproduces the same image counters
        self.o.set_particles(st["pos"] + st["image"] * np.asarray(box), st["vel"], st["mass"], None, type, state, resid)
        self.tabs = []

    # every call below is mirrored on both sides
    def add_table(self, x, e, f, interp=1):
        a = self.e.add_table(x, e, f, interp); b = self.o.add_table(x, e, f, interp)
        assert a == b
        return a

    def nb_tab(self, pairs, tab, rc):
        a = self.e.add_nonbonded("Tabulated"); b = self.o.add_nonbonded(1)
        assert a == b
        for t1, t2 in pairs:
            self.e.nb_set_tabulated(a, t1, t2, tab, rc); self.o.nb_set_tab(b, t1, t2, tab, rc)
        return a

    def nb_lj(self, pairs, eps, sig, rc):
        a = self.e.add_nonbonded("LennardJones"); b = self.o.add_nonbonded(2)
        assert a == b
        for t1, t2 in pairs:
            self.e.nb_set_lj(a, t1, t2, eps, sig, rc, 1); self.o.nb_set_lj(b, t1, t2, eps, sig, rc, 1)
        return a

    def exclusions(self, ex):
        self.e.set_exclusions(ex); self.o.set_exclusions(ex)

    def add_list(self, arity, ids):
        a = self.e.add_list(arity); b = self.o.add_list(arity)
        assert a == b
        if len(ids):
            self.e.list_add(a, ids); self.o.list_add(b, ids)
        return a

    def add_bonded(self, lst, typed=0):
        a = self.e.add_bonded(lst, typed); b = self.o.add_bonded(lst, typed)
        assert a == b
        return a

    def bonded_pot(self, inter, types, kind, params=(), table=-1):
        from chemlab_b200.engine import POT
        self.e.bonded_set_potential(inter, types, kind, params, table)
        self.o.bonded_set_potential(inter, types, POT[kind], params, table)

    def both(self, name, *a, **k):
        ra = getattr(self.e, name)(*a, **k)
        rb = getattr(self.o, name)(*a, **k)
        return ra, rb

    def close(self):
        self.e.close()


def type_pairs(nt):
    return [(a, b) for a in range(nt) for b in range(a, nt)]
