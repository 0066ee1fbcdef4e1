"""GPU parity of ChemicalReaction::React: candidate sets, acceptance, partner resolution, bond lists,
type/state changes, neighbour-property changes, generated angles/dihedrals and exclusions must equal the oracle's
BIT-EXACTLY (integer work), given the shared counter-based per-pair draws or acceptance probability >= 1."""
import numpy as np
import pytest

import clb_testutil as util

pytestmark = pytest.mark.gpu


def _sorted_rows(a):
    a = np.asarray(a, np.int64).reshape(len(a), -1)
    if len(a) == 0:
        return a
    return a[np.lexsort(a.T[::-1])]


def _canon_tuples(a):
    """tuples are direction-free: (i,j,k) == (k,j,i)"""
    a = np.asarray(a, np.int64)
    if len(a) == 0:
        return a
    flip = a[:, 0] > a[:, -1]
    a = a.copy()
    a[flip] = a[flip, ::-1]
    return _sorted_rows(a)


def _reactive_pair(n_side=12, seed=2, nearest=1, p_rate=1e6, interval=10, intramolecular=1, intraresidual=0, max_per_interval=0,
                   steps_before=30, nt_tab=3):
    m = util.melt(n_side, seed=seed)
    n = len(m["pos"])
    v = np.random.default_rng(seed + 7).normal(0, 1, (n, 3))
    state = np.where(m["type"] == 0, 1, 0).astype(np.int32)
    P = util.Pair(m["pos"], m["box"], m["type"], vel=v, state=state, resid=m["resid"], seed=99)
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    r, e, f = util.lj_table()
    tab = P.add_table(r, e, f, 1)
    rl = P.add_list(2, np.zeros((0, 2), np.int64))
    irl = P.add_bonded(rl); P.bonded_pot(irl, (), "Harmonic", (30.0, 0.97))
    P.nb_tab(util.type_pairs(nt_tab), tab, 2.5)
    bl = P.add_list(2, m["bonds"]); al = P.add_list(3, m["angles"])
    ql = P.add_list(4, np.zeros((0, 4), np.int64))
    ib = P.add_bonded(bl); P.bonded_pot(ib, (), "Harmonic", (30.0, 0.97))
    ia = P.add_bonded(al); P.bonded_pot(ia, (), "AngularHarmonic", (1.25, np.pi))
    iq = P.add_bonded(ql); P.bonded_pot(iq, (), "DihedralHarmonic", (0.5, 0.0))
    P.both("set_dt", 0.004); P.both("set_langevin", 1, 1.0, 1.0)
    P.both("reaction_general", 1, interval, nearest, max_per_interval)
    P.both("exclusions_observe", rl); P.both("exclusions_observe", al); P.both("exclusions_observe", ql)
    P.both("topology_observe", bl); P.both("topology_observe", rl)
    for lst, types in ((al, (1, 0, 0)), (al, (1, 0, 2)), (al, (0, 2, 2)), (al, (3, 0, 2)), (al, (0, 2, 3)), (ql, (1, 0, 0, 1)), (ql, (0, 1, 0, 0)),
                       (ql, (0, 1, 0, 2)), (ql, (3, 0, 2, 3)), (ql, (4, 3, 0, 2))):
        P.both("topology_register", lst, types)
    P.both("topology_initialize")
    h = dict(rl=rl, bl=bl, al=al, ql=ql)
    if steps_before:
        P.e.reaction_general(0, interval, nearest, max_per_interval); P.o.reaction_general(0, interval, nearest, max_per_interval)
        P.both("run", steps_before)
        # keep both sides on IDENTICAL coordinates for the set comparisons
        st = P.e.get_particles(fields=("pos", "image", "vel"))
        P.o.set_positions(st["pos"]); P.o.set_velocities(st["vel"])
        P.both("reaction_general", 1, interval, nearest, max_per_interval)
    return m, P, h


def _compare_state(P, h):
    a = P.e.get_particles(fields=("type", "state", "mass")); b = P.o.get()
    assert (a["type"] == b["type"]).all() and (a["state"] == b["state"]).all()
    assert np.allclose(a["mass"], b["mass"], rtol=1e-7)
    assert (_sorted_rows(P.e.list_get(h["rl"], 2)) == _sorted_rows(P.o.list_get(h["rl"], 2))).all()
    assert (_canon_tuples(P.e.list_get(h["al"], 3)) == _canon_tuples(P.o.list_get(h["al"], 3))).all()
    qa, qb = _canon_tuples(P.e.list_get(h["ql"], 4)), _canon_tuples(P.o.list_get(h["ql"], 4))
    assert qa.shape == qb.shape and (qa == qb).all()
    assert (_sorted_rows(P.e.get_exclusions()) == _sorted_rows(P.o.get_exclusions())).all()


def _add_both(P, *a, **k):
    ra = P.e.add_reaction(*a, **k); rb = P.o.add_reaction(*a, **k)
    assert ra == rb
    return ra


@pytest.mark.parametrize("nearest", [1, 0])
def test_reaction_pass_p1_sets_are_bit_exact(nearest):
    m, P, h = _reactive_pair(nearest=nearest)
    _add_both(P, 0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.2, h["rl"], intramolecular=1, intraresidual=0)
    na, nb = P.e.react_now(), P.o.react()
    ca, da = P.e.last_candidates(); cb, db = P.o.candidates()
    assert len(ca) == len(cb) > 50 and (ca == cb).all()
    assert np.allclose(da, db, rtol=1e-12)
    assert na == nb > 20
    _compare_state(P, h)
    # second pass: reacted ends (state 2) are out of the window, leftovers may still react
    assert P.e.react_now() == P.o.react()
    _compare_state(P, h)
    assert (P.e.reaction_counters(1) == [P.o.reaction_counter(0)]).all()
    P.close()


def test_acceptance_draws_match():
    m, P, h = _reactive_pair()
    # p = rate*dt*interval = 0.35
    _add_both(P, 0, 0, 1, 1, 1, 2, 1, 2, 0.35 / (0.004 * 10), 1.3, h["rl"], intramolecular=1, intraresidual=0)
    na, nb = P.e.react_now(), P.o.react()
    ca, _ = P.e.last_candidates(); cb, _ = P.o.candidates()
    assert (ca == cb).all()
    frac = ca[:, 3].mean()
    assert 0.2 < frac < 0.5 and na == nb
    _compare_state(P, h)
    P.close()


def test_type_changes_neighbour_changes_and_generated_tuples():
    # ATRP-like rules (examples/atrp_lj/atrp.cfg): reactant B changes type, its neighbours one/two bonds away change
    m, P, h = _reactive_pair(seed=5)
    r = _add_both(P, 0, 0, 1, 0, 1, 2, 1, 2, 1e6, 1.25, h["rl"], intramolecular=1, intraresidual=0)
    for side, lvl, old, new, kw in ((2, 0, 0, 2, dict(new_mass=1.5)), (3, 1, 1, 3, dict(state_mode=1, state_value=1)),
                                    (3, 2, 1, 3, dict(state_mode=1, state_value=1)), (2, 2, 0, 4, dict(new_q=0.25, state_mode=2, state_value=1))):
        P.e.reaction_add_change(r, side, lvl, old, new, **kw); P.o.reaction_add_change(r, side, lvl, old, new, **kw)
    na, nb = P.e.react_now(), P.o.react()
    assert na == nb > 20
    _compare_state(P, h)
    t = P.e.get_particles(fields=("type",))["type"]
    assert (t == 2).sum() == na and (t == 3).sum() > 0 and (t == 4).sum() > 0
    assert P.e.list_size(h["al"]) > len(m["angles"])       # topology manager generated new angles
    # forces with the new topology still agree
    P.e.compute_forces(); P.o.compute_forces()
    err = util.rel_force_err(P.e.get_particles(fields=("force",))["force"], P.o.get()["force"])
    assert err < 1e-6, err
    P.close()


def test_intramolecular_and_residue_constraints():
    m, P, h = _reactive_pair(seed=6)
    _add_both(P, 0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.3, h["rl"], intramolecular=0, intraresidual=1)
    for _ in range(3):
        assert P.e.react_now() == P.o.react()
        _compare_state(P, h)
    P.close()


def test_two_reactions_and_max_per_interval():
    m, P, h = _reactive_pair(seed=8, max_per_interval=37)
    _add_both(P, 0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.15, h["rl"], intramolecular=1, intraresidual=0)
    _add_both(P, 0, 1, 1, 1, 1, 3, 0, 1, 1e6, 1.1, h["rl"], intramolecular=1, intraresidual=0)
    na, nb = P.e.react_now(), P.o.react()
    assert na == nb == 37
    _compare_state(P, h)
    P.close()


def test_reactive_run_bond_topology_matches():
    # chain_growth_catalytic-style: p >= 1, nearest partner -> RNG-free; 60 steps with a pass every 20
    m, P, h = _reactive_pair(seed=3, interval=20, steps_before=0)
    _add_both(P, 0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.05, h["rl"], intramolecular=1, intraresidual=0)
    P.both("run", 60)
    assert P.e.list_size(h["rl"]) == P.o.list_size(h["rl"]) > 10
    _compare_state(P, h)
    a = P.e.get_particles(); b = P.o.get()
    dx = np.abs((a["pos"] + a["image"] * m["box"]) - (b["pos"] + b["image"] * m["box"])).max()
    assert dx < 5e-4, dx
    t, c = P.e.timers()
    assert c["reaction_passes"] == 3
    P.close()


def test_mixed_tabulated_follows_conversion():
    m, P, h = _reactive_pair(seed=4, nt_tab=2)
    r, e, f = util.lj_table()
    t2 = P.add_table(r, 0.25 * e, 0.25 * f, 1)
    a = P.e.add_nonbonded("MixedTabulated"); b = P.o.add_nonbonded(3)
    n_a = int((m["type"] == 0).sum())
    # type pair (2,2) appears through the reaction; mixing value = N(type 2)/n_a
    P.e.nb_set_mixed(a, 2, 2, 0, t2, 0.0, 2, n_a, 2.5); P.o.nb_set_mixed(b, 2, 2, 0, t2, 0.0, 2, n_a, 2.5)
    rr = _add_both(P, 0, 0, 0, 0, 1, 2, 1, 2, 1e6, 1.2, h["rl"], intramolecular=1, intraresidual=0)
    P.e.reaction_add_change(rr, 3, 0, 0, 2); P.o.reaction_add_change(rr, 3, 0, 0, 2)
    assert P.e.react_now() == P.o.react() > 0
    assert P.e.count_type(2) == P.o.count_type(2) > 0
    P.e.compute_forces(); P.o.compute_forces()
    err = util.rel_force_err(P.e.get_particles(fields=("force",))["force"], P.o.get()["force"])
    assert err < 1e-6, err
    assert abs(P.e.energy(a) - P.o.energy(b)) <= 1e-8 * abs(P.o.energy(b))
    P.close()


def test_atrp_activator_pass_matches_oracle():
    """ATRPActivator (reaction_post_process.py:380-426, examples/atrp_lj/atrp.cfg:16-26) on the device vs the oracle's sequential
    restatement: same selected particles, same activation / deactivation draws, same states / types / masses and catalyst ratios."""
    m, P, h = _reactive_pair(seed=12, steps_before=0)
    # type 0, state 1 = dormant chain ends (flag "A"); type 0, state 3 = active ends (flag "DA") that go back to the dormant state
    P.both("atrp_configure", 150, 0.6, 0.4, 0.5, 0.9, 0.7)
    P.both("atrp_add_center", 0, 1, False, 0, 1.25, float("nan"), 2)
    P.both("atrp_add_center", 0, 3, True, 0, 1.0, float("nan"), -2)
    seen_deact = 0
    for k in range(6):
        (ca, ra), (cb, rb) = P.both("atrp_now")
        assert ca == cb and ra == rb, (k, ca, cb, ra, rb)
        assert ca[0] + ca[1] > 0
        seen_deact += ca[1]
        sa = P.e.get_particles(fields=("type", "state", "mass")); sb = P.o.get()
        assert (sa["type"] == sb["type"]).all() and (sa["state"] == sb["state"]).all() and (sa["mass"] == sb["mass"]).all()
        P.both("run", 3)          # a new step number: new random keys
    assert seen_deact > 0
    assert (P.e.get_particles(fields=("state",))["state"] == 3).sum() > 0
    P.close()
