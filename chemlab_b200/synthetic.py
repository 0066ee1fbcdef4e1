"""Synthetic workloads of BASELINE.json `configs` (no dataset can be fetched here): seeded, vectorised."""
import os

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def lj_table(rmax=3.0, dr=0.002, eps=1.0, sig=1.0, rc=2.5):
    """LJ 12-6, shifted to zero at rc, on the grid of examples/atrp_activator/table_MA_MA.pot
    (dr = 0.002 from r = dr; 1500 rows to r = 3.0) -- SURVEY 8d, config 2."""
    r = dr * np.arange(1, int(round(rmax / dr)) + 1)
    sr6 = (sig / r) ** 6
    src6 = (sig / rc) ** 6
    e = 4 * eps * (sr6 * sr6 - sr6) - 4 * eps * (src6 * src6 - src6)
    f = 24 * eps * (2 * sr6 * sr6 - sr6) / r
    return r, e, f


def trimer_melt(n_side, rho=0.8442, seed=12345, jitter=0.08, kT=1.0, vel_seed=12346):
    """n_side^3 beads on a jittered simple-cubic lattice grouped into A-L-A trimers along x (like the
    MA-ML-MA trimers of examples/atrp_lj/conf.gro); beads that do not fill a trimer stay free A monomers.

    types: 0 = A (reactive end, state 1), 1 = L (middle, state 0).  ids are 0..n-1."""
    rng = np.random.default_rng(seed)
    n = n_side ** 3
    L = (n / rho) ** (1.0 / 3.0)
    a = L / n_side
    g = np.arange(n_side)
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    x = x.ravel(); y = y.ravel(); z = z.ravel()
    pos = (np.stack([x, y, z], 1) + 0.5) * a + rng.uniform(-jitter, jitter, (n, 3)) * a
    ntri = n_side // 3
    in_tri = x < 3 * ntri
    k = x % 3
    idx = np.arange(n, dtype=np.int64)
    type_ = np.where(in_tri & (k == 1), 1, 0).astype(np.int32)
    state = np.where(type_ == 0, 1, 0).astype(np.int32)
    first = idx[in_tri & (k == 0)]
    bonds = np.concatenate([np.stack([first, first + 1], 1), np.stack([first + 1, first + 2], 1)])
    angles = np.stack([first, first + 1, first + 2], 1)
    resid = np.zeros(n, np.int32)
    tri_id = np.arange(len(first), dtype=np.int32)
    resid[first] = tri_id; resid[first + 1] = tri_id; resid[first + 2] = tri_id
    free = idx[~in_tri]
    resid[free] = len(first) + np.arange(len(free), dtype=np.int32)
    vel = np.random.default_rng(vel_seed).normal(0.0, np.sqrt(kT), (n, 3))
    vel -= vel.mean(0)
    excl = np.concatenate([bonds, angles[:, [0, 2]]])
    return dict(n=n, box=np.array([L, L, L]), pos=pos, vel=vel, type=type_, state=state, resid=resid, ids=idx,
                mass=np.ones(n), bonds=bonds, angles=angles, exclusions=excl)


def replicated_melt(n_side, tile="melt_tile_20.npz"):
    """The C2 melt as a periodic replication of an EQUILIBRATED 8000-bead tile ("a pre-equilibrated
    replicated tile of a small box"; the tile was equilibrated for 30,000 steps by tests/golden/make_melt_tile.py).
    n_side must be a multiple of the tile's 20 beads per edge: 100 -> 5x5x5 tiles = 1,000,000 beads
    (320,000 A-L-A trimers + 40,000 free A monomers).  Same keys as trimer_melt()."""
    d = np.load(os.path.join(DATA, tile))
    ts = int(d["n_side"])
    if n_side % ts:
        raise ValueError("n_side must be a multiple of %d" % ts)
    k = n_side // ts
    nt = len(d["pos"])
    tb = d["box"]
    gz, gy, gx = np.meshgrid(np.arange(k), np.arange(k), np.arange(k), indexing="ij")
    shift = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1) * tb            # (k^3, 3)
    nrep = len(shift)
    pos = (d["pos"][None, :, :] + shift[:, None, :]).reshape(-1, 3)
    vel = np.tile(d["vel"], (nrep, 1))
    off = (np.arange(nrep, dtype=np.int64) * nt)

    def rep(a):
        a = np.asarray(a, np.int64)
        return (a[None, :, :] + off[:, None, None]).reshape(-1, a.shape[1])
    nres = int(d["resid"].max()) + 1
    resid = (d["resid"][None, :].astype(np.int64) + (np.arange(nrep) * nres)[:, None]).reshape(-1).astype(np.int32)
    n = nrep * nt
    return dict(n=n, box=tb * k, pos=pos, vel=vel, type=np.tile(d["type"], nrep).astype(np.int32), state=np.tile(d["state"], nrep).astype(np.int32),
                resid=resid, ids=np.arange(n, dtype=np.int64), mass=np.ones(n), bonds=rep(d["bonds"]), angles=rep(d["angles"]),
                exclusions=rep(d["exclusions"]))


def setup_reactive_melt(api, sysd, rc=2.5, dt=0.005, kT=1.0, gamma=1.0, interval=200, p_accept=0.05, reactions=True,
                        cutoff_react=1.2):
    """Wire config 2 onto any object with the Engine method names: tabulated LJ pairs, harmonic bonds K=30 r0=0.97, harmonic angle 180 deg K=1.25
    (= GROMACS 2.5 halved, gromacs_topology.py:1073), step-growth reaction A(1,2)+A(1,2)->A(1):A(1),
    new L-A-A angles through the topology manager, new bonds excluded."""
    r, e, f = lj_table(rc=rc)
    tab = api.add_table(r, e, f, 1)
    rl = api.add_list(2)                      # reaction bonds: chem_fpl_<group> comes first (reaction_setup.py:467)
    irl = api.add_bonded(rl, 0)
    api.bonded_set_potential(irl, (), 1, (30.0, 0.97))
    nb = api.add_nonbonded(1)
    for a, b in ((0, 0), (0, 1), (1, 1)):
        api.nb_set_tabulated(nb, a, b, tab, rc)
    bl = api.add_list(2); api.list_add(bl, sysd["bonds"])
    ib = api.add_bonded(bl, 0); api.bonded_set_potential(ib, (), 1, (30.0, 0.97))
    al = api.add_list(3); api.list_add(al, sysd["angles"])
    ia = api.add_bonded(al, 0); api.bonded_set_potential(ia, (), 3, (1.25, np.pi))
    api.set_exclusions(sysd["exclusions"])
    api.set_dt(dt)
    api.set_langevin(1, kT, gamma)
    handles = dict(nb=nb, bonds=ib, angles=ia, react_bonds=irl, react_list=rl, bond_list=bl, angle_list=al,
                   # generic description used by bench.py (state hand-over, parity, observables)
                   lists=dict(react=(rl, 2), bonds=(bl, 2), angles=(al, 3)),
                   initial=dict(react=0, bonds=len(sysd["bonds"]), angles=len(sysd["angles"])),
                   energies=dict(nb=nb, bonds=ib, angles=ia, react_bonds=irl), table_bytes=3 * len(r) * 8)
    if reactions:
        rate = p_accept / (dt * interval)
        api.reaction_general(0, interval, 1, 0)
        handles["reaction"] = api.add_reaction(0, 0, 1, 1, 1, 2, 1, 2, rate, cutoff_react, rl, intramolecular=1, intraresidual=0)
        api.exclusions_observe(rl); api.exclusions_observe(al)
        api.topology_observe(bl); api.topology_observe(rl)
        api.topology_register(al, (1, 0, 0))
        api.topology_initialize()
    return handles


# ======================================================================================================================
# Configs 3-5 of BASELINE.json: the reference's shipped multi-table systems, replicated to the benchmark size.
#
# The base system (4000 / 4000 / 800 beads) is set up ONCE by the real chemlab driver (start_simulation.py on the
# espressopp surface) against a recording Engine: every C-ABI call the driver would make is kept on a tape.  The tape is
# then replayed on a real engine with the particle, tuple and exclusion arrays replicated k x k x k (ids offset per
# replica, box scaled by k, conversion totals scaled by k^3).  So the benchmark runs exactly the force field, lists and
# reactions the driver builds from the shipped .top / .cfg / params files -- without Python loops over millions of atoms.
EXAMPLES = {
    # name: (example directory name, default replications per edge, extra driver arguments)
    "c3": ("hyperbranched", 10, ("--rng_seed", "5", "--gen_velocity", "True")),                              # 4000 x 1000 = 4 M beads
    "c4": ("dacron", 8, ("--rng_seed", "7", "--t_hybrid_bond", "0", "--gen_velocity", "True")),              # 4000 x 512  = 2.05 M beads (hybrid bonds dropped: SURVEY 8d)
    "c5": ("rim135", 27, ("--rng_seed", "11", "--start_ar", "0", "--gen_velocity", "True")),                 # 800 x 19683 = 15.7 M beads
}


def prepare_example(src_dir, dst_dir, example):
    """Copy a shipped example into a scratch directory, unpack its `.pot` tables (tests/golden/*/tables.npz) and synthesise the
    angle / dihedral tables the reference does not ship (.MISSING_LARGE_BLOBS): smooth stand-ins on the usual grids."""
    import shutil
    shutil.copytree(src_dir, dst_dir)
    th = np.radians(np.arange(0.5, 180.01, 0.5))
    ph = np.radians(np.arange(-180.0, 180.01, 1.0))

    def write(name, x, e, f):
        with open(os.path.join(dst_dir, name), "w") as fh:
            fh.writelines("%15.8g %15.8g %15.8g\n" % r for r in zip(x, e, f))
    if example == "hyperbranched":
        for k in range(11):
            t0, K = np.radians(100 + 6 * k), 40.0 + 3 * k
            write("table_a%d.pot" % k, th, 0.5 * K * (th - t0) ** 2, -K * (th - t0))
        for k in range(8):
            write("table_d%d.pot" % k, ph, 2.0 * (1 + np.cos(2 * ph - 0.3 * k)), 4.0 * np.sin(2 * ph - 0.3 * k))
    if example == "dacron_restrict":
        # examples/dacron/restrict: the includes, the exclusion list and every table are byte-identical to no_water/test_1 and come
        # from that fixture.  Dropped as out of scope, exactly like there: the dummy-water type Z with its
        # func-11 (dynamic-resolution) rows -- only the unused ReleaseMolecule extension would create Z particles.
        base = os.path.join(os.path.dirname(src_dir), "dacron")
        for f in ("diol_cg.itp", "ter_cg.itp", "exclusion_topol.list", "tables.npz"):
            shutil.copy(os.path.join(base, f), os.path.join(dst_dir, f))
        top = os.path.join(dst_dir, "topol.top")
        lines = open(top).read().split("\n")
        with open(top, "w") as fh:
            fh.write("\n".join((";" + l) if (l.split()[:1] == ["Z"] or l.split()[1:2] == ["Z"]) else l for l in lines))
    if example in ("dacron", "dacron_restrict"):
        for k in range(2):
            write("table_d%d.pot" % k, ph, 1.5 * (1 + np.cos(3 * ph - 0.4 * k)), 4.5 * np.sin(3 * ph - 0.4 * k))
    npz = os.path.join(dst_dir, "tables.npz")
    if os.path.exists(npz):
        with np.load(npz) as z:
            for name in z.files:
                with open(os.path.join(dst_dir, name + ".pot"), "w") as fh:
                    fh.writelines("%15.8g %15.8g %15.8g\n" % tuple(r) for r in z[name])
    return dst_dir


class RecordingEngine:
    """Stands in for chemlab_b200.Engine while the driver sets a system up: records every call, answers queries from the
    recorded state, runs nothing.  Handles are the sequential ones the real engine hands out (checked again at replay)."""
    ID_CALLS = ("set_particles", "list_add", "set_exclusions")

    def __init__(self, box, rc_max, skin, seed=0, device=0):
        self.box = np.array(box, float)
        self.rc, self.skin, self.seed = float(rc_max), float(skin), int(seed)
        self.tape = []
        self.ntables = self.nlists = self.ninters = self.nreactions = 0
        self.list_arity, self.list_rows = {}, {}
        self.inter_kind = {}
        self.P = None
        self.excl = np.zeros((0, 2), np.int64)
        self._step = 0
        self.n = 0

    def _rec(self, name, *a, **k):
        self.tape.append((name, a, k))

    # -- calls that return handles
    def add_table(self, x, e, f, interp=1):
        self._rec("add_table", np.array(x, float), np.array(e, float), np.array(f, float), int(interp))
        self.ntables += 1
        return self.ntables - 1

    def add_list(self, arity):
        self._rec("add_list", int(arity))
        self.list_arity[self.nlists] = int(arity); self.list_rows[self.nlists] = []
        self.nlists += 1
        return self.nlists - 1

    def add_nonbonded(self, kind):
        self._rec("add_nonbonded", kind)
        self.inter_kind[self.ninters] = "nb"
        self.ninters += 1
        return self.ninters - 1

    def add_bonded(self, lst, typed=0):
        self._rec("add_bonded", int(lst), int(typed))
        self.inter_kind[self.ninters] = "bonded"
        self.ninters += 1
        return self.ninters - 1

    def add_reaction(self, *a, **k):
        self._rec("add_reaction", *a, **k)
        self.nreactions += 1
        return self.nreactions - 1

    # -- calls that carry particle ids
    def set_particles(self, ids, type, pos, mass, vel=None, q=None, state=None, res_id=None):
        n = len(ids)
        z = np.zeros
        self.P = dict(ids=np.array(ids, np.int64), type=np.array(type, np.int32), pos=np.array(pos, float), mass=np.array(mass, float),
                      vel=np.array(vel, float) if vel is not None else z((n, 3)), q=np.array(q, float) if q is not None else z(n),
                      state=np.array(state, np.int32) if state is not None else z(n, np.int32),
                      res_id=np.array(res_id, np.int32) if res_id is not None else z(n, np.int32))
        self.n = n
        self._rec("set_particles")

    def list_add(self, lst, ids):
        ids = np.array(ids, np.int64)
        if ids.size:
            ids = ids.reshape(-1, self.list_arity[lst])
            self.list_rows[lst].append(ids)
            self._rec("list_add", int(lst), ids)

    def set_exclusions(self, pairs):
        self.excl = np.array(pairs, np.int64).reshape(-1, 2)
        self._rec("set_exclusions", self.excl)

    # -- queries
    def num_particles(self):
        return self.n

    def get_particles(self, ids=None, fields=("pos", "vel", "force", "type", "state", "mass", "image", "q", "res_id"), out=None):
        sel = slice(None) if ids is None else (np.searchsorted(self.P["ids"], np.asarray(ids, np.int64)))
        m = self.n if ids is None else len(ids)
        src = dict(self.P, force=np.zeros((self.n, 3)), image=np.zeros((self.n, 3), np.int32))
        return {f: src[f][sel] for f in fields}

    def list_size(self, lst):
        return int(sum(len(r) for r in self.list_rows[lst]))

    def list_get(self, lst, arity):
        r = self.list_rows[lst]
        return np.concatenate(r) if r else np.zeros((0, arity), np.int64)

    def get_exclusions(self):
        return self.excl

    def energy(self, inter):
        return 0.0

    def kinetics(self):
        return np.zeros(3)

    def count_type(self, type, state=-1):
        t = self.P["type"] == type
        if state >= 0:
            t &= self.P["state"] == state
        return int(t.sum())

    def step(self):
        return self._step

    def run(self, n):
        self._step += int(n)

    run_continue = run

    def timers(self):
        return ({k: 0.0 for k in ("pair", "bonded", "neighbour", "integrate", "comm", "reaction", "other", "total")},
                {k: 0 for k in ("steps", "rebuilds", "launches", "list_entries", "reaction_passes", "reaction_events", "ghosts", "interacting_pairs")})

    def reaction_counters(self, n):
        return np.zeros(n, np.int64)

    def react_now(self):
        return 0

    def close(self):
        pass

    def get_option(self, name):
        return 0.0

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)

        def rec(*a, **k):           # every other call is a plain setter: keep it verbatim
            self._rec(name, *a, **k)
        return rec


def record_example(example_src, example, extra_args=()):
    """Run the chemlab driver on the shipped example against a RecordingEngine; returns the recorder (tape + base state)."""
    import contextlib
    import io
    import tempfile
    from .espressopp import _context as C
    from . import start_simulation as S
    tmp = tempfile.mkdtemp(prefix="clb_%s_" % example)
    d = prepare_example(example_src, os.path.join(tmp, example), example)
    cwd = os.getcwd()
    real = C.Engine
    C.Engine = RecordingEngine
    os.chdir(d)
    try:
        params = {}
        for line in open("params"):
            if "=" in line and not line.strip().startswith("#"):
                k, v = line.strip().split("=", 1); params[k] = v
        one_iter = str(min(int(params.get("int_step", 1000)), int(params.get("trj_collect", 10 ** 9)) or 10 ** 9))
        with contextlib.redirect_stdout(io.StringIO()):
            r = S.main(["@params", "--run", one_iter] + list(extra_args))
        rec = r["system"]._ctx.engine
    finally:
        C.Engine = real
        os.chdir(cwd)
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return rec


class TapeWorkload:
    """One of BASELINE.json configs 3-5: a recorded base system replicated k^3 times (see the section comment above)."""
    def __init__(self, name, example_root, scale=0):
        ex, k_default, extra = EXAMPLES[name]
        self.name, self.example = name, ex
        self.k = int(scale) if scale else k_default
        self.rec = record_example(os.path.join(example_root, ex), ex, extra)
        r = self.rec
        self.rc, self.skin = r.rc, r.skin
        self.base_n = r.n
        self.n = r.n * self.k ** 3
        self.rho = r.n / float(np.prod(r.box))
        g = {nm: (a, kw) for nm, a, kw in r.tape if nm in ("set_dt", "set_langevin", "reaction_general")}
        self.dt = float(g["set_dt"][0][0])
        self.kT, self.gamma = float(g["set_langevin"][0][1]), float(g["set_langevin"][0][2])
        rg = [a for nm, a, kw in r.tape if nm == "reaction_general" and a[0]]
        self.interval = int(rg[-1][1]) if rg else 1
        self.nearest = bool(rg[-1][2]) if rg else True
        self.ntables_pair = len({(a[3]) for nm, a, kw in r.tape if nm == "nb_set_tabulated"} | {a[3] for nm, a, kw in r.tape if nm == "nb_set_mixed"}
                                | {a[4] for nm, a, kw in r.tape if nm == "nb_set_mixed"})
        self.description = ("%s: examples/%s as shipped (%d beads, %d pair tables, driver-built force field and reactions) replicated %dx%dx%d = %d beads"
                            % (name.upper(), ex, r.n, self.ntables_pair, self.k, self.k, self.k, self.n))

    def l_half(self):
        return (2.0 * np.pi / 3.0) * (self.rc + self.skin) ** 3 * self.rho

    def config(self, gpus):
        return {"workload": self.description, "n_beads": self.n, "rho": self.rho, "rc": self.rc, "skin": self.skin, "dt": self.dt, "kT": self.kT,
                "gamma": self.gamma, "reaction_interval": self.interval, "nearest": self.nearest, "pair_tables": self.ntables_pair,
                "l2": "per-step working set exceeds the 126 MB L2; no explicit flush", "parallelism": "slab%d" % gpus if gpus > 1 else "single"}

    def _rep_ids(self, a):
        """Tuple / pair rows of the base system -> rows of all k^3 replicas.  Ids are dense (replica r adds r*n).  The shipped
        coordinates are folded into the box, so a bonded partner may sit across the periodic boundary: member m of a tuple then
        belongs to the NEIGHBOUR replica in that direction (replica index shifted by the minimum-image vector relative to the
        first member), which keeps the replicated system exactly equivalent to the periodic base system."""
        a = np.asarray(a, np.int64)
        k, n, P = self.k, self.base_n, self.rec.P
        ids0 = int(P["ids"][0])
        idx = a - ids0
        pos, L = P["pos"], self.rec.box
        p0 = pos[idx[:, 0]]
        g = np.arange(k)
        gz, gy, gx = np.meshgrid(g, g, g, indexing="ij")
        G = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1)                       # (k^3, 3), replica r = (gz*k + gy)*k + gx
        out = np.empty((k ** 3, len(a), a.shape[1]), np.int64)
        for m in range(a.shape[1]):
            nm = -np.rint((pos[idx[:, m]] - p0) / L).astype(np.int64)               # (rows, 3) image of member m relative to member 0
            T = (G[:, None, :] + nm[None, :, :]) % k                                # (k^3, rows, 3)
            r2 = (T[:, :, 2] * k + T[:, :, 1]) * k + T[:, :, 0]
            out[:, :, m] = ids0 + idx[None, :, m] + r2 * n
        return out.reshape(-1, a.shape[1])

    def system(self):
        P, k = self.rec.P, self.k
        gz, gy, gx = np.meshgrid(np.arange(k), np.arange(k), np.arange(k), indexing="ij")
        shift = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1) * self.rec.box
        nrep = k ** 3
        nres = int(P["res_id"].max()) + 1
        ids0 = int(P["ids"][0])
        if not np.array_equal(P["ids"], ids0 + np.arange(self.base_n)):
            raise RuntimeError("replication needs dense particle ids")
        n = self.n
        return dict(n=n, box=self.rec.box * k, ids=ids0 + np.arange(n, dtype=np.int64),
                    pos=(P["pos"][None, :, :] + shift[:, None, :]).reshape(-1, 3), vel=np.tile(P["vel"], (nrep, 1)),
                    type=np.tile(P["type"], nrep), state=np.tile(P["state"], nrep), mass=np.tile(P["mass"], nrep), q=np.tile(P["q"], nrep),
                    resid=(P["res_id"][None, :].astype(np.int64) + (np.arange(nrep) * nres)[:, None]).reshape(-1).astype(np.int32),
                    exclusions=self._rep_ids(self.rec.excl) if len(self.rec.excl) else np.zeros((0, 2), np.int64))

    def setup(self, api, sysd):
        """Replay the tape (everything but engine creation and set_particles) on `api`."""
        k3 = self.k ** 3
        lists, inters, initial = {}, {}, {}
        nt = nl = ni = nr = 0
        for name, a, kw in self.rec.tape:
            if name in ("set_particles", "join"):      # (a tape recorded under torchrun holds the recorder's own join)
                continue
            if name == "add_table":
                h = api.add_table(*a); assert h == nt; nt += 1
            elif name == "add_list":
                h = api.add_list(*a); assert h == nl; lists["list%d" % h] = (h, a[0]); initial["list%d" % h] = 0; nl += 1
            elif name == "add_nonbonded":
                h = api.add_nonbonded(*a); assert h == ni; inters["nb%d" % h] = h; ni += 1
            elif name == "add_bonded":
                h = api.add_bonded(*a); assert h == ni; inters["bonded%d_list%d" % (h, a[0])] = h; ni += 1
            elif name == "add_reaction":
                h = api.add_reaction(*a, **kw); assert h == nr; nr += 1
            elif name == "list_add":
                rows = self._rep_ids(a[1])
                api.list_add(a[0], rows); initial["list%d" % a[0]] += len(rows)
            elif name == "set_exclusions":
                api.set_exclusions(self._rep_ids(a[0]) if len(a[0]) else a[0])
            elif name == "nb_set_mixed":
                a = list(a); a[7] = a[7] * k3           # conversion total: N(type) / total keeps its meaning
                api.nb_set_mixed(*a)
            elif name == "set_velocities":
                api.set_velocities(np.tile(np.asarray(a[0], float).reshape(-1, 3), (k3, 1)))
            elif name in ("run", "run_continue", "close", "set_option"):
                continue
            else:
                getattr(api, name)(*a, **kw)
        react_lists = sorted({a[10] if len(a) > 10 else kw.get("lst") for nm, a, kw in self.rec.tape if nm == "add_reaction"})
        nb = [h for nm, h in inters.items() if nm.startswith("nb")]
        tb = sum(a[0].size * 3 * 8 for nm, a, kw in self.rec.tape if nm == "add_table")
        return dict(lists=lists, initial=initial, energies=inters, nb=nb[0], react_lists=react_lists, table_bytes=tb)


def make_workload(name, scale=0, example_root=None):
    if example_root is None:
        example_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    return TapeWorkload(name, example_root, scale)
