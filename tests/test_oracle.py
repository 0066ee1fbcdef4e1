"""CPU checks that pin the oracle itself: RNG known answers, force = -grad(energy), cell list = brute force."""
import numpy as np
import pytest

from oracle import pyoracle
import clb_testutil as util


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert pyoracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert pyoracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert pyoracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _small(seed=3, n_side=9):
    m = util.melt(n_side, seed=seed)
    n = len(m["pos"])
    o = pyoracle.Oracle(n, m["box"], 2.5, 0.3, seed=7)
    o.set_particles(m["pos"], None, np.ones(n), None, m["type"], np.ones(n, np.int32), m["resid"])
    return m, o, n


def test_cell_list_equals_brute_force():
    m, o, n = _small()
    o.set_exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    a = o.pairs()
    b = o.pairs_brute()
    b = b[np.lexsort((b[:, 1], b[:, 0]))]
    assert len(a) == len(b) and (a == b).all()
    # exclusions really removed
    ex = set(map(tuple, util.exclusions_from(m["bonds"], m["angles"])))
    assert not (ex & set(map(tuple, a)))


def _fd_check(o, inters, n, h=1e-6, probes=(0, 5, 17, 100), tol=2e-5):
    o.compute_forces()
    st = o.get()
    x0, f0 = st["pos"].copy(), st["force"].copy()
    for i in probes:
        for c in range(3):
            e = []
            for s in (+1, -1):
                x = x0.copy(); x[i, c] += s * h
                o.set_positions(x); o.rebuild(); o.compute_forces()
                e.append(sum(o.energy(k) for k in inters))
            fd = -(e[0] - e[1]) / (2 * h)
            assert abs(fd - f0[i, c]) <= tol * max(1.0, abs(f0[i, c])), (i, c, fd, f0[i, c])
    o.set_positions(x0)


def test_pair_forces_are_energy_gradient():
    m, o, n = _small()
    r, e, f = util.lj_table()
    # energy/force consistency needs a smooth table: use cubic interpolation for the FD check only
    tab = o.add_table(r, e, f, 1)
    nb = o.add_nonbonded(1)
    for a in (0, 1):
        for b in (a, 1):
            o.nb_set_tab(nb, a, b, tab, 2.5)
    lj = pyoracle.Oracle(n, m["box"], 2.5, 0.3)
    lj.set_particles(m["pos"], None, np.ones(n), None, m["type"])
    k = lj.add_nonbonded(2)
    for a in (0, 1):
        for b in (a, 1):
            lj.nb_set_lj(k, a, b, 1.0, 1.0, 2.5, 1)
    _fd_check(lj, [k], n)
    # tabulated LJ reproduces analytic LJ to the table's interpolation error
    o.compute_forces(); lj.compute_forces()
    assert util.rel_force_err(o.get()["force"], lj.get()["force"]) < 2e-4
    assert abs(o.energy(nb) - lj.energy(k)) < 1e-4 * abs(lj.energy(k))


def test_bonded_forces_are_energy_gradient():
    m, o, n = _small()
    rng = np.random.default_rng(0)
    bl = o.add_list(2); o.list_add(bl, m["bonds"])
    al = o.add_list(3); o.list_add(al, m["angles"])
    quads = np.array([(i, i + 1, i + 2, i + 3) for i in range(0, 200, 9)], np.int64)
    ql = o.add_list(4); o.list_add(ql, quads)
    ib = o.add_bonded(bl); o.bonded_set_potential(ib, (), 1, (30.0, 0.97))
    ia = o.add_bonded(al); o.bonded_set_potential(ia, (), 3, (1.25, 2.6))
    iq = o.add_bonded(ql); o.bonded_set_potential(iq, (), 8, (2.0, 0.7))
    # tabulated angle + dihedral (cubic so that f = -de/dx holds between knots)
    th = np.linspace(0.0, np.pi, 361)
    ta = o.add_table(th, 3.0 * (th - 2.0) ** 2, -6.0 * (th - 2.0), 3)
    ph = np.linspace(-np.pi, np.pi, 721)
    td = o.add_table(ph, 1.5 * (1 + np.cos(2 * ph - 0.4)), 3.0 * np.sin(2 * ph - 0.4), 3)
    al2 = o.add_list(3); o.list_add(al2, m["angles"][:50])
    ia2 = o.add_bonded(al2); o.bonded_set_potential(ia2, (), 4, (), ta)
    ql2 = o.add_list(4); o.list_add(ql2, quads)
    iq2 = o.add_bonded(ql2); o.bonded_set_potential(iq2, (), 5, (), td)
    # natural-spline end conditions limit the tabulated-angle consistency near theta = pi (straight trimers)
    _fd_check(o, [ib, ia, iq, ia2, iq2], n, probes=(0, 1, 2, 3, 10, 11), tol=2e-3)
    o2 = pyoracle.Oracle(n, m["box"], 2.5, 0.3, seed=7)
    o2.set_particles(m["pos"], None, np.ones(n), None, m["type"])
    for ar, ids, kind, par in ((2, m["bonds"], 1, (30.0, 0.97)), (3, m["angles"], 3, (1.25, 2.6)), (4, quads, 8, (2.0, 0.7)),
                               (2, m["bonds"][:40], 7, (30.0, 0.0, 1.5)), (2, m["bonds"][40:90], 9, (30.0, 0.0, 1.5, 1.0, 1.0)),
                               # 1-4 pairs: LJ on a pair list (kind 10) with cutoff 5 (everything inside) and a shift
                               (2, np.array([(a[0], a[2]) for a in m["angles"]], np.int64), 10, (0.8, 1.1, 5.0, 0.01))):
        l = o2.add_list(ar); o2.list_add(l, ids)
        o2.bonded_set_potential(o2.add_bonded(l), (), kind, par)
    _fd_check(o2, [0, 1, 2, 3, 4, 5], n, probes=(0, 1, 2, 3, 10, 11), tol=1e-6)
    f = o.get()["force"]
    assert np.abs(f.sum(0)).max() < 1e-9  # Newton's third law


def test_akima_and_linear_tables():
    o = pyoracle.Oracle(1, [10, 10, 10], 1.0, 0.1)
    x = np.linspace(0.1, 2.0, 96)
    y = np.sin(3 * x)
    t1 = o.add_table(x, y, -3 * np.cos(3 * x), 1)
    t2 = o.add_table(x, y, -3 * np.cos(3 * x), 2)
    for xv in (0.1, 0.55, 1.234, 1.99):
        e1, f1, _ = o.table_eval(t1, xv)
        e2, f2, _ = o.table_eval(t2, xv)
        i = int((xv - 0.1) / (x[1] - x[0])); i = min(i, 94)
        b = (xv - x[i]) / (x[1] - x[0])
        assert abs(e1 - ((1 - b) * y[i] + b * y[i + 1])) < 1e-12
        assert abs(e2 - np.sin(3 * xv)) < 2e-5       # Akima interpolates a smooth function to O(h^3..4)
        assert abs(f2 + 3 * np.cos(3 * xv)) < 2e-4
    assert o.table_eval(t1, 2.5)[2] == 1 and o.table_eval(t1, 0.05)[2] == 1  # out of range flagged


def test_akima_and_cubic_match_scipy_on_a_shipped_table():
    """U13 (itype 2 / 3): the oracle's Akima and natural cubic spline agree with scipy's public implementations on a table
    shipped by the reference (examples/hyperbranched/table_b0.pot) -- energy and force columns splined independently."""
    import os
    from scipy.interpolate import Akima1DInterpolator, CubicSpline
    d = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "table_b0.pot"))
    x, e, f = d[:, 0], d[:, 1], d[:, 2]
    o = pyoracle.Oracle(8, [10, 10, 10], 2.5, 0.3)
    ta, tc = o.add_table(x, e, f, 2), o.add_table(x, e, f, 3)
    xs = np.random.default_rng(0).uniform(x[3], x[-4], 500)       # interior: the end treatment of Akima differs between codes
    oa = np.array([o.table_eval(ta, v)[:2] for v in xs])
    oc = np.array([o.table_eval(tc, v)[:2] for v in xs])
    for got, ref in ((oa[:, 0], Akima1DInterpolator(x, e)(xs)), (oa[:, 1], Akima1DInterpolator(x, f)(xs)),
                     (oc[:, 0], CubicSpline(x, e, bc_type="natural")(xs)), (oc[:, 1], CubicSpline(x, f, bc_type="natural")(xs))):
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


def test_tabulated_pair_forces_match_independent_numpy_brute_force():
    """SURVEY 3.4 restated a second time, in numpy, O(N^2), minimum image: idx = floor((r-r0)/dr), F = (1-b) f[idx] + b f[idx+1],
    f_i += F/r * d for r <= rc, pairs in the exclusion set skipped.  The C oracle (cell list + Verlet list) must agree."""
    import clb_testutil as util
    m = util.melt(7, seed=11)
    n = len(m["pos"]); box = m["box"]; rc = 2.5
    o = pyoracle.Oracle(n, box, rc, 0.3, seed=1)
    o.set_particles(m["pos"], np.zeros((n, 3)), np.ones(n), None, m["type"], None, m["resid"])
    ex = util.exclusions_from(m["bonds"], m["angles"])
    o.set_exclusions(ex)
    r, e, f = util.lj_table()
    tab = o.add_table(r, e, f, 1)
    nb = o.add_nonbonded(1)
    for a, b in util.type_pairs(2):
        o.nb_set_tab(nb, a, b, tab, rc)
    o.compute_forces()
    got = o.get()["force"]
    d = m["pos"][:, None, :] - m["pos"][None, :, :]
    d -= box * np.rint(d / box)
    rr = np.sqrt((d * d).sum(-1))
    mask = (rr <= rc) & ~np.eye(n, dtype=bool)
    mask[ex[:, 0], ex[:, 1]] = False; mask[ex[:, 1], ex[:, 0]] = False
    dr = r[1] - r[0]
    s = (rr - r[0]) / dr
    idx = np.clip(np.floor(s).astype(int), 0, len(r) - 2)
    bfr = s - idx
    F = (1 - bfr) * f[idx] + bfr * f[idx + 1]
    E = (1 - bfr) * e[idx] + bfr * e[idx + 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(mask, F / rr, 0.0)
    ref = (w[:, :, None] * d).sum(1)
    assert util.rel_force_err(got, ref) < 1e-12
    assert abs(o.energy(nb) - 0.5 * E[mask].sum()) <= 1e-12 * abs(0.5 * E[mask].sum())


def test_react_matches_an_independent_python_restatement():
    """ChemicalReaction::React restated a second time in plain Python on the oracle's own pair
    list and Philox draws: candidate rows, acceptance, UniqueA -> UniqueB (nearest and random), one reaction per particle,
    max_per_interval, state deltas and the bond list must equal what the C oracle does."""
    STREAM_REACT, STREAM_PARTNER = 0x52454143, 0x50415254
    # last case: reaction 0 is a RestrictReaction (reaction_setup.py:74-75,115-126) whose connectivity map names every third
    # Verlet pair (given in either order, with duplicates) -- a pair outside the map is no candidate of that reaction
    for nearest, cap, restrict in ((1, 0, 0), (0, 0, 0), (1, 25, 0), (0, 0, 1)):
        m = util.melt(10, seed=21)
        n = len(m["pos"]); box = m["box"]; seed = 77; dt = 0.004; interval = 10
        state = np.where(m["type"] == 0, 1, 0).astype(np.int32)
        o = pyoracle.Oracle(n, box, 2.5, 0.3, seed=seed)
        o.set_particles(m["pos"], np.zeros((n, 3)), np.ones(n), None, m["type"], state, m["resid"])
        o.set_exclusions(util.exclusions_from(m["bonds"], m["angles"]))
        rl = o.add_list(2)
        o.set_dt(dt); o.reaction_general(1, interval, nearest, cap)
        # two reactions: A(1,2)+A(1,2)->A(1):A(1) with p = 0.6, and A(1,2)+L(0,1)->A(1):L(1) with p >= 1
        specs = [dict(t1=0, t2=0, d1=1, d2=1, w1=(1, 2), w2=(1, 2), rate=0.6 / (dt * interval), rc=1.25),
                 dict(t1=0, t2=1, d1=1, d2=1, w1=(1, 2), w2=(0, 1), rate=1e6, rc=1.1)]
        for s in specs:
            o.add_reaction(s["t1"], s["t2"], s["d1"], s["d2"], s["w1"][0], s["w1"][1], s["w2"][0], s["w2"][1], s["rate"], s["rc"], rl,
                           intramolecular=1, intraresidual=0)
        pairs = o.pairs()
        if restrict:
            cmap = pairs[::3]
            specs[0]["conn"] = {(int(min(a, b)), int(max(a, b))) for a, b in cmap}
            o.reaction_define_connections(0, np.concatenate([cmap[:, ::-1], cmap[:50]]))
        typ, st, resid, pos = m["type"].copy(), state.copy(), m["resid"], m["pos"]
        step = o.step()

        def draw(stream, i, j, r):
            return pyoracle.philox([i, j, step & 0xffffffff, (((step >> 32) & 0xffffffff) << 8) ^ r], [seed & 0xffffffff, ((seed >> 32) & 0xffffffff) ^ stream])

        def side_ok(s, a, b):
            return typ[a] == s["t1"] and typ[b] == s["t2"] and s["w1"][0] <= st[a] < s["w1"][1] and s["w2"][0] <= st[b] < s["w2"][1]
        cands = []
        for i, j in pairs:
            d = pos[i] - pos[j]; d -= box * np.rint(d / box); d2 = float(d @ d)
            for r, s in enumerate(specs):
                if side_ok(s, i, j): a, b = i, j
                elif side_ok(s, j, i): a, b = j, i
                else: continue
                if resid[a] == resid[b] or not (d2 < s["rc"] ** 2): continue
                if "conn" in s and (int(min(i, j)), int(max(i, j))) not in s["conn"]: continue
                w = draw(STREAM_REACT, int(i), int(j), r); h = draw(STREAM_PARTNER, int(i), int(j), r)
                cands.append(dict(a=int(a), b=int(b), r=r, d2=d2, acc=(w[0] / 4294967296.0) < s["rate"] * dt * interval, rnd=(h[0] << 32) | h[1]))
        cands.sort(key=lambda c: (c["a"], c["b"], c["r"]))
        key = (lambda c, partner: ((c["d2"] if nearest else c["rnd"]), c[partner], c["r"]))
        alive = [c for c in cands if c["acc"]]
        best = {}
        for c in alive:
            if c["a"] not in best or key(c, "b") < key(best[c["a"]], "b"): best[c["a"]] = c
        alive = [c for c in alive if best[c["a"]] is c]
        best = {}
        for c in alive:
            if c["b"] not in best or key(c, "a") < key(best[c["b"]], "a"): best[c["b"]] = c
        alive = [c for c in alive if best[c["b"]] is c]
        used, events = set(), []
        for c in alive:
            if c["a"] in used or c["b"] in used or (cap and len(events) >= cap): continue
            used.update((c["a"], c["b"])); events.append(c)
        for c in events:
            st[c["a"]] += specs[c["r"]]["d1"]; st[c["b"]] += specs[c["r"]]["d2"]
        nev = o.react()
        rows, d2o = o.candidates()
        assert len(rows) == len(cands) > 100
        if not restrict:
            n_unrestricted = sum(c["r"] == 0 for c in cands)
        else:
            assert 0 < sum(c["r"] == 0 for c in cands) < 0.5 * n_unrestricted
            assert all((min(c["a"], c["b"]), max(c["a"], c["b"])) in specs[0]["conn"] for c in cands if c["r"] == 0)
        assert (rows == np.array([[c["a"], c["b"], c["r"], int(c["acc"])] for c in cands])).all()
        assert np.allclose(d2o, [c["d2"] for c in cands], rtol=1e-13)
        assert nev == len(events) > (10 if not cap else 0) and (not cap or nev == cap)
        got = o.list_get(rl, 2)
        want = np.array([[c["a"], c["b"]] for c in events])
        assert (got[np.lexsort((got[:, 1], got[:, 0]))] == want[np.lexsort((want[:, 1], want[:, 0]))]).all()
        assert (o.get()["state"] == st).all()


def test_atrp_pass_matches_an_independent_numpy_restatement():
    """orc_atrp_now against a numpy restatement written from the description in REFERENCE_UNVERIFIED.md (U22): candidates by
    (type, state), the num_particles smallest (Philox word 0, index) keys, reaction when the uniform from word 1 is below
    k * ratio of the start of the pass, catalyst ratios moved by delta * (n_act - n_deact) / num_particles."""
    rng = np.random.default_rng(3)
    n, seed, ATRP = 600, 77, 0x41545250
    box = np.array([12.0, 12.0, 12.0])
    typ = rng.integers(0, 3, n).astype(np.int32); st = rng.integers(0, 3, n).astype(np.int32)
    o = pyoracle.Oracle(n, box, 2.5, 0.3, seed=seed)
    o.set_particles(rng.random((n, 3)) * box, np.zeros((n, 3)), np.ones(n), None, typ, st, np.arange(n, dtype=np.int32))
    num, ra, rd, delta, ka, kd = 40, 0.7, 0.3, 0.8, 0.9, 0.6
    o.atrp_configure(num, ra, rd, delta, ka, kd)
    centres = [(0, 1, False, 2, 1.5, 1), (2, 2, True, 0, 1.0, -1), (1, 0, False, -1, -1.0, 2)]
    for t, s, de, nt, nm, ds in centres:
        o.atrp_add_center(t, s, de, nt, nm, float("nan"), ds)
    typ = typ.copy(); st = st.copy(); mass = np.ones(n)
    for step in (0, 0, 0):           # the oracle's step counter stays 0: the same keys, a changing candidate set
        cen = np.full(n, -1)
        for k in reversed(range(len(centres))):
            cen[(typ == centres[k][0]) & (st == centres[k][1])] = k
        idx = np.nonzero(cen >= 0)[0]
        words = np.array([pyoracle.philox([int(i), 0, step, 0], [seed & 0xffffffff, (seed >> 32) ^ ATRP]) for i in idx], dtype=np.uint64).reshape(-1, 4)
        order = np.lexsort((idx, words[:, 0]))[:num]
        na = nd = 0
        for j in order:
            i, k = idx[j], cen[idx[j]]
            t, s, de, nt, nm, ds = centres[k]
            if (float(words[j, 1]) + 0.5) / 4294967296.0 < (kd * rd if de else ka * ra):
                st[i] += ds
                if nt >= 0:
                    typ[i] = nt
                if nm > 0:
                    mass[i] = nm
                na += (not de); nd += de
        d = delta * (na - nd) / num
        ra, rd = min(1.0, max(0.0, ra - d)), min(1.0, max(0.0, rd + d))
        (ca, cd), (oa, od) = o.atrp_now()
        g = o.get()
        assert (ca, cd) == (na, nd) and na + nd > 0
        assert abs(oa - ra) < 1e-15 and abs(od - rd) < 1e-15
        assert (g["type"] == typ).all() and (g["state"] == st).all() and (g["mass"] == mass).all()


def test_velocity_verlet_and_langevin_match_an_independent_numpy_restatement():
    """VelocityVerlet::run + LangevinThermostat restated in numpy for non-interacting particles plus harmonic
    bonds evaluated by the restatement itself: run-entry heat-up kick (sqrt 3, its own stream), half kick, drift, fold with image
    counters, force, friction + noise  -gamma m v + sqrt(24 kT gamma / dt) sqrt(m) (u - 1/2)  with u = (word + 1/2) / 2^32 from
    Philox(seed ^ stream; particle, 0, step), thermostat restricted to a type list, second half kick.  Two consecutive runs (the
    step counter keys the draws of the second)."""
    LANG, HEAT = 0x4c414e47, 0x48454154
    rng = np.random.default_rng(11)
    n, seed, dt, kT, gamma = 60, 4242, 0.004, 1.3, 0.7
    box = np.array([6.0, 7.0, 8.0])
    pos = rng.uniform(0, 1, (n, 3)) * box
    vel = rng.normal(0, 1, (n, 3))
    mass = rng.uniform(0.5, 3.0, n)
    typ = rng.integers(0, 3, n).astype(np.int32)
    bonds = np.array([[2 * k, 2 * k + 1] for k in range(n // 2)], np.int64)
    pos[bonds[:, 1]] = pos[bonds[:, 0]] + rng.normal(0, 0.4, (n // 2, 3))           # partners close by, some across the box faces
    K, r0 = 17.0, 0.8
    o = pyoracle.Oracle(n, box, 2.5, 0.3, seed=seed)
    o.set_particles(pos, vel, mass, None, typ, None, None)
    bl = o.add_list(2); o.list_add(bl, bonds)
    ib = o.add_bonded(bl, 0); o.bonded_set_potential(ib, (), 1, (K, r0))
    o.set_dt(dt); o.set_langevin(1, kT, gamma, types=(0, 2))                        # type 1 is not thermalised

    x, v = pos.copy(), vel.copy()
    img = np.zeros((n, 3), np.int64)
    im = np.floor(x / box); x -= im * box; img += im.astype(np.int64)               # set_particles folds

    def bonded(x):
        f = np.zeros((n, 3))
        d = x[bonds[:, 0]] - x[bonds[:, 1]]; d -= box * np.rint(d / box)
        r = np.linalg.norm(d, axis=1)
        fr = (-2.0 * K * (r - r0) / r)[:, None] * d                                  # U = K (r - r0)^2
        np.add.at(f, bonds[:, 0], fr); np.add.at(f, bonds[:, 1], -fr)
        return f

    def thermo(f, v, stream, step, scale):
        pref2 = np.sqrt(24.0 * kT * gamma / dt) * scale
        for i in range(n):
            if typ[i] == 1:
                continue
            w = pyoracle.philox([i, 0, step & 0xffffffff, step >> 32], [seed & 0xffffffff, (seed >> 32) ^ stream])
            u = (np.array(w[:3], float) + 0.5) / 4294967296.0
            f[i] += -gamma * mass[i] * v[i] + pref2 * np.sqrt(mass[i]) * (u - 0.5)

    step = 0
    for nsteps in (7, 5):
        f = bonded(x); thermo(f, v, HEAT, step, np.sqrt(3.0))
        for it in range(nsteps):
            v += (0.5 * dt / mass)[:, None] * f
            x += dt * v
            f = bonded(x); thermo(f, v, LANG, step + it, 1.0)
            v += (0.5 * dt / mass)[:, None] * f
        step += nsteps
        o.run(nsteps)
        g = o.get()
        assert o.step() == step
        assert np.abs(g["vel"] - v).max() < 1e-12
        # the oracle folds at rebuilds only: stored position + image * L is the unfolded coordinate at any time
        assert np.abs(g["pos"] + g["image"] * box - (x + img * box)).max() < 1e-12


def test_topology_manager_tuples_match_path_enumeration_on_the_final_graph():
    """TopologyManager after a reaction pass against a formulation that shares nothing with the oracle's
    per-event emission: the new angles (dihedrals) must be exactly the simple paths of 3 (4) particles in the FINAL bond graph
    that run through at least one bond created in this pass and whose FINAL type tuple, read in either direction, is registered --
    each once, in the list of the first matching registration; a DynamicExcludeList observing those lists gains (first, last)
    of every new tuple and (a, b) of every new bond.  The neighbour-property rule (U18) is checked too: a particle exactly one
    bond from the type_1-side reactant whose type matches takes the new type.  Two passes: in the second, ends that already
    carry a reaction bond react again, so paths run through old reaction bonds as well."""
    m = util.melt(9, seed=5)
    n = len(m["pos"]); box = m["box"]
    state = np.where(m["type"] == 0, 1, 0).astype(np.int32)
    o = pyoracle.Oracle(n, box, 2.5, 0.3, seed=3)
    o.set_particles(m["pos"], np.zeros((n, 3)), np.ones(n), None, m["type"], state, m["resid"])
    ex0 = util.exclusions_from(m["bonds"], m["angles"])
    o.set_exclusions(ex0)
    bl = o.add_list(2); o.list_add(bl, m["bonds"])
    rl = o.add_list(2)
    al = o.add_list(3); o.list_add(al, m["angles"])
    al2 = o.add_list(3)
    ql = o.add_list(4)
    o.set_dt(0.004); o.reaction_general(1, 10, 1, 0)
    for lst in (rl, al, al2, ql):
        o.excl_observe(lst)
    o.tm_observe(bl); o.tm_observe(rl)
    # types: 0 = A (end), 1 = L (middle), 3 = middle bead whose neighbour reacted as type_1
    regs3 = [(al2, (3, 0, 0)), (al, (3, 0, 0)), (al, (1, 0, 0)), (al2, (0, 0, 0))]            # (al, (3,0,0)) is shadowed by the entry before it
    regs4 = [(ql, (3, 0, 0, 3)), (ql, (3, 0, 0, 1)), (ql, (1, 0, 0, 1)), (ql, (0, 1, 0, 0)), (ql, (0, 3, 0, 0)), (ql, (0, 0, 0, 0)),
             (ql, (3, 0, 0, 0))]                                                             # (1,0,0,0) is left unregistered
    for lst, t in regs3 + regs4:
        o.tm_register(lst, t)
    o.tm_initialize()
    r = o.add_reaction(0, 0, 1, 1, 1, 3, 1, 3, 1e6, 1.25, rl, intramolecular=1, intraresidual=0)      # every end may bind twice
    o.reaction_add_change(r, 1, 1, 1, 3)                     # neighbours of the type_1 reactant: L -> 3
    canon = lambda t: tuple(t) if t[0] < t[-1] else tuple(t[::-1])
    have = {al: {canon(t) for t in m["angles"].tolist()}, al2: set(), ql: set()}
    old_b = np.zeros((0, 2), np.int64)
    want_t = m["type"].copy()
    want_ex = {tuple(p) for p in ex0.tolist()}
    seen_kinds = set()
    for pas in range(2):
        nev = o.react()
        allb = o.list_get(rl, 2).astype(np.int64)
        newb = allb[len(old_b):]
        assert nev == len(newb) and nev > (30 if pas == 0 else 3)
        old_b = allb
        typ = o.get()["type"]
        adj = {i: set() for i in range(n)}
        for a, b in np.concatenate([m["bonds"], allb]).tolist():
            adj[a].add(b); adj[b].add(a)
        for a in newb[:, 0].tolist():
            for q in adj[a]:
                if want_t[q] == 1:
                    want_t[q] = 3
        assert (typ == want_t).all()
        new_edges = {(min(a, b), max(a, b)) for a, b in newb.tolist()}
        is_new = lambda a, b: (min(a, b), max(a, b)) in new_edges

        def first_match(regs, ids):
            ty = tuple(int(typ[i]) for i in ids)
            for lst, t in regs:
                if t == ty or t == ty[::-1]:
                    seen_kinds.add(t)
                    return lst
            return None
        want = {al: set(), al2: set(), ql: set()}
        for j in range(n):
            for i in adj[j]:
                for k in adj[j]:
                    if i < k and (is_new(i, j) or is_new(j, k)):
                        lst = first_match(regs3, (i, j, k))
                        if lst is not None:
                            want[lst].add(canon((i, j, k)))
        for j in range(n):
            for k in adj[j]:
                if j < k:
                    for i in adj[j] - {k}:
                        for l in adj[k] - {j, i}:
                            if is_new(i, j) or is_new(j, k) or is_new(k, l):
                                lst = first_match(regs4, (i, j, k, l))
                                if lst is not None:
                                    want[lst].add(canon((i, j, k, l)))
        want_ex |= new_edges
        for lst, ar in ((al, 3), (al2, 3), (ql, 4)):
            rows = [canon(t) for t in o.list_get(lst, ar).tolist()]
            assert len(rows) == len(set(rows)), "a tuple was emitted twice"
            got_new = set(rows) - have[lst]
            assert len(rows) == len(have[lst]) + len(got_new)
            assert want[lst].isdisjoint(have[lst])           # a path through a bond of this pass cannot have existed before
            assert got_new == want[lst], (pas, lst, len(got_new), len(want[lst]))
            have[lst] |= got_new
            want_ex |= {(min(t[0], t[-1]), max(t[0], t[-1])) for t in got_new}
        got_ex = {(min(a, b), max(a, b)) for a, b in o.get_exclusions().tolist()}
        assert got_ex == want_ex
    assert len(have[al]) > len(m["angles"]) + 10 and len(have[al2]) > 10 and len(have[ql]) > 10
    assert {(3, 0, 0), (1, 0, 0), (0, 0, 0), (0, 0, 0, 0)} <= seen_kinds          # the second pass produced paths through old reaction bonds


@pytest.mark.parametrize("criterion", [0, 1])
def test_verlet_list_stays_complete_between_rebuilds(criterion):
    """Skin/2 resort rule: at any step of a Langevin run the Verlet list in use -- built at the last rebuild with
    radius rc + skin -- must contain every non-excluded pair that is now within rc.  Checked with numpy O(N^2) distances at
    several points of a hot melt run, for the reference's criterion (sum of per-step maxima, 0) and the true-displacement one (1);
    and the rule must not be trivially satisfied by rebuilding at every step."""
    m = util.melt(9, seed=8)
    n = len(m["pos"]); box = m["box"]; rc = 2.5
    o = pyoracle.Oracle(n, box, rc, 0.3, seed=5)
    rng = np.random.default_rng(2)
    o.set_particles(m["pos"], rng.normal(0, 1.2, (n, 3)), np.ones(n), None, m["type"], None, m["resid"])
    ex = util.exclusions_from(m["bonds"], m["angles"])
    o.set_exclusions(ex)
    r, e, f = util.lj_table()
    tab = o.add_table(r, e, f, 1)
    nb = o.add_nonbonded(1)
    for t1, t2 in util.type_pairs(2):
        o.nb_set_tab(nb, t1, t2, tab, rc)
    bl = o.add_list(2); o.list_add(bl, m["bonds"])
    ib = o.add_bonded(bl, 0); o.bonded_set_potential(ib, (), 1, (30.0, 0.97))
    o.set_option("resort_criterion", criterion)
    o.set_dt(0.004); o.set_langevin(1, 1.5, 1.0)
    exs = {tuple(p) for p in ex.tolist()}
    iu = np.triu_indices(n, 1)
    steps = 0
    for chunk in (3, 4, 5, 7, 9, 11):
        o.run(chunk); steps += chunk
        x = o.get()["pos"]
        d = x[iu[0]] - x[iu[1]]; d -= box * np.rint(d / box)
        close = (d * d).sum(1) <= rc * rc
        need = {(int(a), int(b)) for a, b in zip(iu[0][close], iu[1][close])} - exs
        have = {tuple(p) for p in o.pairs().tolist()}
        assert need <= have, (steps, len(need - have))
    assert 1 <= o.nrebuild() < steps // 2


def test_cell_list_equals_numpy_brute_force_on_random_boxes():
    """Randomised: non-cubic boxes from barely 2 cut-offs wide to many cells, densities from dilute to dense, particles given
    outside the box (several images away), random exclusions.  The Verlet pair set must equal an O(N^2) numpy minimum-image
    search with the inclusive rule r^2 <= (rc + skin)^2 (U1), and the image counters must fold every particle into [0, L)."""
    rng = np.random.default_rng(123)
    for trial in range(12):
        rc, skin = float(rng.uniform(0.8, 2.5)), float(rng.uniform(0.05, 0.5))
        box = (rc + skin) * rng.uniform(2.05, 6.0, 3)
        n = int(rng.integers(30, 400))
        pos = rng.uniform(-2.0, 3.0, (n, 3)) * box                       # up to two images below, three above
        o = pyoracle.Oracle(n, box, rc, skin, seed=1)
        o.set_particles(pos, None, np.ones(n), None, np.zeros(n, np.int32), None, None)
        iu = np.triu_indices(n, 1)
        ex = np.column_stack(iu)[rng.random(len(iu[0])) < 0.02]
        o.set_exclusions(ex)
        got = {tuple(p) for p in o.pairs().tolist()}
        d = pos[iu[0]] - pos[iu[1]]; d -= box * np.rint(d / box)
        close = (d * d).sum(1) <= (rc + skin) ** 2
        want = {(int(a), int(b)) for a, b in zip(iu[0][close], iu[1][close])} - {tuple(p) for p in ex.tolist()}
        assert got == want, (trial, len(got), len(want), box, rc, skin)
        g = o.get()
        assert (g["pos"] >= 0).all() and (g["pos"] < box).all()
        assert np.abs(g["pos"] + g["image"] * box - pos).max() < 1e-9
