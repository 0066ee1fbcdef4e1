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
