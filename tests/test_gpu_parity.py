"""GPU parity tests: the CUDA engine (through the C-ABI) against the fp64 CPU oracle on identical inputs.

Bars (BASELINE.json north_star): neighbour pair sets, reaction candidate sets and bond lists bit-exact;
per-particle forces <= 1e-6 relative; energies <= 1e-8 relative.  The relative force error is
max_i |f_i - fref_i| / max(|fref_i|, rms|fref|)  (clb_testutil.rel_force_err).
"""
import numpy as np
import pytest

import clb_testutil as util

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-6
ENERGY_TOL = 1e-8


def _melt_pair(n_side=12, seed=1, rc=2.5, skin=0.3, rho=0.8442, vel=False, **kw):
    m = util.melt(n_side, rho=rho, seed=seed)
    n = len(m["pos"])
    v = None
    if vel:
        v = np.random.default_rng(seed + 100).normal(0, 1.0, (n, 3))
        v -= v.mean(0)
    # shift everything so that particles straddle the periodic boundary and image counters matter
    pos = m["pos"] + np.array([0.37, -1.3, 2.9]) * m["box"]
    P = util.Pair(pos, m["box"], m["type"], vel=v, state=np.ones(n, np.int32), resid=m["resid"], rc=rc, skin=skin, **kw)
    return m, P


def _setup_tab_lj(P, rc=2.5, nt=2):
    r, e, f = util.lj_table(rc=rc)
    tab = P.add_table(r, e, f, 1)
    return P.nb_tab(util.type_pairs(nt), tab, rc)


def _compare_forces(P, inters, tol=FORCE_TOL):
    P.e.compute_forces(); P.o.compute_forces()
    fe = P.e.get_particles(fields=("force",))["force"]
    fo = P.o.get()["force"]
    err = util.rel_force_err(fe, fo)
    assert err <= tol, "force mismatch %.3e" % err
    for k in inters:
        a, b = P.e.energy(k), P.o.energy(k)
        assert abs(a - b) <= ENERGY_TOL * max(abs(b), 1e-300), ("energy", k, a, b)
    return err


def test_state_roundtrip_and_lattice():
    m, P = _melt_pair(9)
    st = P.e.get_particles()
    L = m["box"][0]
    want = (m["pos"] + np.array([0.37, -1.3, 2.9]) * m["box"])
    assert np.abs((st["pos"] + st["image"] * L) - want).max() <= L / 2 ** 32   # lattice resolution
    assert (st["type"] == m["type"]).all() and (st["state"] == 1).all() and (st["res_id"] == m["resid"]).all()
    o = P.o.get()
    assert np.abs(o["pos"] - st["pos"]).max() < 1e-12 and (o["image"] == st["image"]).all()
    P.close()


@pytest.mark.parametrize("n_side,rc,skin", [(9, 2.5, 0.3), (12, 2.5, 0.3), (14, 1.2, 0.4), (20, 2.5, 0.3)])
def test_verlet_pair_set_bit_exact(n_side, rc, skin):
    m, P = _melt_pair(n_side, rc=rc, skin=skin)
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    a = P.e.pairs()
    b = P.o.pairs().astype(np.int64)
    assert len(a) == len(b) and (a == b).all()
    if n_side <= 12:
        c = P.o.pairs_brute().astype(np.int64)
        c = c[np.lexsort((c[:, 1], c[:, 0]))]
        assert (a == c).all()
    P.close()


def test_pair_set_with_arbitrary_ids_and_block_sizes():
    m = util.melt(10, seed=5)
    n = len(m["pos"])
    ids = np.random.default_rng(3).permutation(np.arange(1000, 1000 + 3 * n, 3))[:n]   # sparse, shuffled ids
    P = util.Pair(m["pos"], m["box"], m["type"], ids=ids)
    want = P.ids[P.o.pairs()]            # oracle index k <-> k-th smallest id
    for bx in (1, 3, 8, 16):
        P.e.set_option("block_cells", bx)
        got = P.e.pairs()
        assert len(got) == len(want) and (got == want).all(), bx
    P.close()


def test_tabulated_pair_forces_and_energy():
    m, P = _melt_pair(14)
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    nb = _setup_tab_lj(P)
    err = _compare_forces(P, [nb])
    # tables in global memory instead of shared memory: same numbers
    P.e.set_option("tables_in_smem", 0); P.e.decompose()
    err2 = _compare_forces(P, [nb])
    assert abs(err - err2) < 1e-9
    P.close()


def test_lj_and_mixed_kind_pair_forces():
    m, P = _melt_pair(12, seed=4)
    r, e, f = util.lj_table()
    tab = P.add_table(r, e, f, 1)
    k1 = P.nb_tab([(0, 0)], tab, 2.5)
    k2 = P.nb_lj([(0, 1), (1, 1)], 0.8, 1.05, 2.3)
    _compare_forces(P, [k1, k2])
    P.close()


def test_mixed_tabulated():
    m, P = _melt_pair(10, seed=6)
    r, e, f = util.lj_table()
    t1 = P.add_table(r, e, f, 1)
    t2 = P.add_table(r, 0.5 * e, 0.5 * f, 1)
    a = P.e.add_nonbonded("MixedTabulated"); b = P.o.add_nonbonded(3)
    for x, y in util.type_pairs(2):
        P.e.nb_set_mixed(a, x, y, t1, t2, 0.3, -1, 1.0, 2.5)
        P.o.nb_set_mixed(b, x, y, t1, t2, 0.3, -1, 1.0, 2.5)
    _compare_forces(P, [a])
    P.close()


def test_noncubic_box():
    rng = np.random.default_rng(2)
    box = np.array([9.0, 11.5, 14.2])
    n = 1400
    pos = rng.uniform(0, 1, (n, 3)) * box
    # push apart overlapping random points a little: keep r > 0.8 by rejection
    from scipy.spatial import cKDTree
    for _ in range(30):
        t = cKDTree(pos, boxsize=box)
        bad = np.unique(t.query_pairs(0.85, output_type="ndarray")[:, 1])
        if len(bad) == 0:
            break
        pos[bad] = rng.uniform(0, 1, (len(bad), 3)) * box
    ty = (np.arange(n) % 2).astype(np.int32)
    P = util.Pair(pos, box, ty)
    a, b = P.e.pairs(), P.o.pairs().astype(np.int64)
    assert len(a) == len(b) and (a == b).all()
    nb = _setup_tab_lj(P)
    _compare_forces(P, [nb], tol=FORCE_TOL)
    P.close()


def test_bonded_forces_all_kinds():
    m, P = _melt_pair(12, seed=8)
    bl = P.add_list(2, m["bonds"]); al = P.add_list(3, m["angles"])
    quads = np.array([(i, i + 1, i + 2, i + 3) for i in range(0, 600, 5)], np.int64)
    ql = P.add_list(4, quads)
    ib = P.add_bonded(bl); P.bonded_pot(ib, (), "Harmonic", (30.0, 0.97))
    ia = P.add_bonded(al); P.bonded_pot(ia, (), "AngularHarmonic", (1.25, np.pi))
    iq = P.add_bonded(ql); P.bonded_pot(iq, (), "DihedralHarmonic", (2.0, 0.7))
    # tabulated variants on separate lists: bond (linear + Akima), angle, dihedral
    r = np.linspace(0.002, 3.0, 1500)
    tb1 = P.add_table(r, 15 * (r - 1.0) ** 2, -30 * (r - 1.0), 1)
    tb2 = P.add_table(r, 15 * (r - 1.0) ** 2, -30 * (r - 1.0), 2)
    th = np.linspace(0.0, np.pi, 1801)
    ta = P.add_table(th, 3.0 * (th - 2.0) ** 2, -6.0 * (th - 2.0), 1)
    ph = np.linspace(-np.pi, np.pi, 721)
    td = P.add_table(ph, 1.5 * (1 + np.cos(2 * ph - 0.4)), 3.0 * np.sin(2 * ph - 0.4), 2)
    l1 = P.add_list(2, m["bonds"][:300]); i1 = P.add_bonded(l1); P.bonded_pot(i1, (), "Tabulated", (), tb1)
    l2 = P.add_list(2, m["bonds"][300:700]); i2 = P.add_bonded(l2); P.bonded_pot(i2, (), "Tabulated", (), tb2)
    l3 = P.add_list(3, m["angles"][:200]); i3 = P.add_bonded(l3); P.bonded_pot(i3, (), "TabulatedAngular", (), ta)
    l4 = P.add_list(4, quads); i4 = P.add_bonded(l4); P.bonded_pot(i4, (), "TabulatedDihedral", (), td)
    # type-dispatched bonds: (0,1) harmonic, (1,1) not registered -> skipped on both sides
    l5 = P.add_list(2, m["bonds"]); i5 = P.add_bonded(l5, typed=1); P.bonded_pot(i5, (1, 0), "Harmonic", (10.0, 1.1))
    # Kremer-Grest bonds of examples/pccg_lj: FENE (func 7) and FENE + LJ (func 9, gromacs_topology.py:935-961)
    l6 = P.add_list(2, m["bonds"][:250]); i6 = P.add_bonded(l6); P.bonded_pot(i6, (), "FENE", (30.0, 0.0, 1.5))
    l7 = P.add_list(2, m["bonds"][250:]); i7 = P.add_bonded(l7); P.bonded_pot(i7, (), "FENELennardJones", (30.0, 0.0, 1.5, 1.0, 1.0))
    # 1-4 [ pairs ]: Lennard-Jones on a pair list with a cutoff and the 'auto' shift (gromacs_topology.py:1314-1411); the 1-3 pairs of
    # the trimers serve as the pair list (r ~ 1.9: inside the cutoff 2.2 for most, beyond it for the stretched ones)
    p13 = np.array([(a[0], a[2]) for a in m["angles"]], np.int64)
    sr6 = (1.1 / 2.2) ** 6
    l8 = P.add_list(2, p13); i8 = P.add_bonded(l8); P.bonded_pot(i8, (), "LennardJones", (0.8, 1.1, 2.2, 4 * 0.8 * (sr6 * sr6 - sr6)))
    l9 = P.add_list(2, p13); i9 = P.add_bonded(l9, typed=1); P.bonded_pot(i9, (0, 0), "LennardJones", (0.5, 1.0, 1.9, 0.0))
    _compare_forces(P, [ib, ia, iq, i1, i2, i3, i4, i5, i6, i7, i8, i9])
    P.close()


def test_everything_together_forces():
    m, P = _melt_pair(14, seed=9)
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    nb = _setup_tab_lj(P)
    bl = P.add_list(2, m["bonds"]); al = P.add_list(3, m["angles"])
    ib = P.add_bonded(bl); P.bonded_pot(ib, (), "Harmonic", (30.0, 0.97))
    ia = P.add_bonded(al); P.bonded_pot(ia, (), "AngularHarmonic", (1.25, np.pi))
    _compare_forces(P, [nb, ib, ia])
    P.close()


def test_table_range_error_is_fatal():
    from chemlab_b200 import EngineError
    m, P = _melt_pair(9)
    r = np.linspace(1.2, 3.0, 901)        # table starts above the closest contacts
    sr6 = 1.0 / r ** 6
    tab = P.e.add_table(r, 4 * (sr6 * sr6 - sr6), 24 * (2 * sr6 * sr6 - sr6) / r, 1)
    nb = P.e.add_nonbonded("Tabulated")
    P.e.nb_set_tabulated(nb, 0, 0, tab, 2.5); P.e.nb_set_tabulated(nb, 0, 1, tab, 2.5); P.e.nb_set_tabulated(nb, 1, 1, tab, 2.5)
    with pytest.raises(EngineError):
        P.e.compute_forces()
    P.close()


# ------------------------------------------------------------------------------------------ integrator
def _md_pair(n_side=10, seed=3, langevin=True, dt=0.004, crit=1):
    m, P = _melt_pair(n_side, seed=seed, vel=True)
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    _setup_tab_lj(P)
    bl = P.add_list(2, m["bonds"]); al = P.add_list(3, m["angles"])
    ib = P.add_bonded(bl); P.bonded_pot(ib, (), "Harmonic", (30.0, 0.97))
    ia = P.add_bonded(al); P.bonded_pot(ia, (), "AngularHarmonic", (1.25, np.pi))
    P.both("set_dt", dt)
    P.both("set_langevin", int(langevin), 1.0, 1.0)
    P.both("set_option", "resort_criterion", crit)
    return m, P


def _unfolded(st, box):
    return st["pos"] + st["image"] * np.asarray(box)


@pytest.mark.parametrize("langevin", [False, True])
def test_single_step_matches_oracle(langevin):
    m, P = _md_pair(langevin=langevin)
    P.both("run", 1)
    a = P.e.get_particles(); b = P.o.get()
    q = m["box"][0] / 2 ** 32
    # one step: positions agree to the lattice resolution, velocities to fp32 storage rounding
    assert np.abs(_unfolded(a, m["box"]) - _unfolded(b, m["box"])).max() <= 1.01 * q
    assert np.abs(a["vel"] - b["vel"]).max() <= 2.0 ** -22 * max(1.0, np.abs(b["vel"]).max())
    P.close()


@pytest.mark.parametrize("fuse,crit", [(1, 1), (0, 1), (1, 0)])
def test_trajectory_tracks_oracle(fuse, crit):
    m, P = _md_pair(langevin=True, crit=crit)
    P.e.set_option("fuse_integrator", fuse)
    nsteps = 60
    P.both("run", 25)           # two runs: run-entry logic (heat-up recalc) is exercised twice
    P.both("run", nsteps - 25)
    a = P.e.get_particles(); b = P.o.get()
    # fp32 velocity storage + lattice positions: drift from the fp64 trajectory stays ~1e-5 over 60 steps
    dx = np.abs(_unfolded(a, m["box"]) - _unfolded(b, m["box"])).max()
    dv = np.abs(a["vel"] - b["vel"]).max()
    assert dx < 2e-4 and dv < 2e-3, (dx, dv)
    assert P.e.step() == P.o.step() == nsteps
    t, c = P.e.timers()
    assert c["rebuilds"] >= 2 and c["steps"] == nsteps
    if crit == 0:
        assert abs(c["rebuilds"] - P.o.nrebuild()) <= 2
    P.close()


def test_langevin_thermostat_reaches_temperature():
    m, P = _md_pair(n_side=12, langevin=True, dt=0.005)
    P.e.set_langevin(1, 1.5, 2.0)
    P.e.run(1500)
    T = np.mean([P.e.kinetics()[1] for _ in range(1)])
    ts = []
    for _ in range(10):
        P.e.run(50); ts.append(P.e.kinetics()[1])
    assert abs(np.mean(ts) - 1.5) < 0.08, ts
    P.close()


def test_nve_energy_conservation():
    m, P = _md_pair(n_side=12, langevin=False, dt=0.002)
    P.e.run(100)

    def etot():
        return sum(P.e.energy(k) for k in range(3)) + P.e.kinetics()[0]
    e0 = etot()
    P.e.run(400)
    e1 = etot()
    assert abs(e1 - e0) / len(m["pos"]) < 2e-3, (e0, e1)
    P.close()


def test_run_is_bit_reproducible():
    res = []
    for _ in range(2):
        m, P = _md_pair(n_side=9, langevin=True)
        P.e.run(120)
        st = P.e.get_particles()
        res.append((st["pos"].copy(), st["vel"].copy(), st["force"].copy()))
        P.close()
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all() and (res[0][2] == res[1][2]).all()


def test_cell_pair_resort_criterion():
    """resort_criterion=2 (option): when the fastest bead has moved more than skin/2 the lists are kept as long as
    D(c) + D(c') <= skin for all cells within two of each other.  The lists must stay COMPLETE: at every step the forces from
    the current (possibly old) lists equal the oracle's forces on the same positions, and rebuilds are rarer than with the
    global-maximum rule."""
    rebuilds = {}
    for crit in (1, 2):
        m, P = _md_pair(n_side=16, langevin=True, dt=0.005, crit=crit)
        P.e.set_langevin(1, 1.3, 1.0)                   # a warm melt: fast beads, frequent resorts
        assert P.e.get_option("resort_criterion") == crit
        worst = 0.0
        for k in range(40):
            P.e.run_continue(1) if k else P.e.run(1)
            if crit == 2:
                st = P.e.get_particles(fields=("pos",))
                P.e.compute_forces()                                              # lists valid -> no rebuild: pair + bonded forces from the lists as they are
                fe = P.e.get_particles(fields=("force",))["force"]
                P.o.set_positions(st["pos"])
                P.o.compute_forces()
                worst = max(worst, util.rel_force_err(fe, P.o.get()["force"]))
        if crit == 2:
            assert worst < 1e-6, worst
        rebuilds[crit] = P.e.timers()[1]["rebuilds"]
        P.close()
    assert rebuilds[2] <= rebuilds[1], rebuilds


def test_cap_force():
    """integrator.CapForce (src/start_simulation.py:320-324): forces longer than the cap are scaled back to it, on the engine as in
    the oracle; a short run with the cap and the thermostat tracks the oracle."""
    m, P = _md_pair(n_side=10, langevin=True)
    P.both("set_cap_force", 12.0)
    P.e.compute_forces(); P.o.compute_forces()
    fe = P.e.get_particles(fields=("force",))["force"]; fo = P.o.get()["force"]
    assert util.rel_force_err(fe, fo) < 1e-6
    nf = np.linalg.norm(fe, axis=1)
    assert nf.max() <= 12.0 * (1 + 1e-12) and (np.abs(nf - 12.0) < 1e-9).sum() > 10          # some forces were capped
    P.both("run", 30)
    a = P.e.get_particles(); b = P.o.get()
    assert np.abs(_unfolded(a, m["box"]) - _unfolded(b, m["box"])).max() < 2e-4
    P.both("set_cap_force", -1.0)
    P.e.compute_forces(); P.o.compute_forces()
    assert np.linalg.norm(P.e.get_particles(fields=("force",))["force"], axis=1).max() > 12.0
    P.close()
