"""Edge cases of the hot path on the GPU engine vs the oracle: degenerate inputs, every pair-kernel variant, multi-table
systems with missing type pairs, reactions without candidates."""
import numpy as np
import pytest

import clb_testutil as util

pytestmark = pytest.mark.gpu


def _forces(P):
    P.e.compute_forces(); P.o.compute_forces()
    return P.e.get_particles(fields=("force",))["force"], P.o.get()["force"]


def test_free_particles_no_bonds_no_exclusions_no_reactions():
    m = util.melt(10, seed=4, trimers=False)
    n = len(m["pos"])
    v = np.random.default_rng(1).normal(0, 1, (n, 3))
    P = util.Pair(m["pos"], m["box"], m["type"], vel=v)
    r, e, f = util.lj_table()
    tab = P.add_table(r, e, f, 1)
    nb = P.nb_tab([(0, 0)], tab, 2.5)
    a, b = P.e.pairs(), P.o.pairs()
    assert len(a) == len(b) and (a == b).all()
    fe, fo = _forces(P)
    assert util.rel_force_err(fe, fo) < 1e-6
    P.both("set_dt", 0.004); P.both("set_langevin", 0, 1.0, 1.0)
    P.both("run", 25)
    sa, sb = P.e.get_particles(), P.o.get()
    d = sa["pos"] - (sb["pos"] - sb["image"] * m["box"])
    d -= m["box"] * np.rint(d / m["box"])
    assert np.abs(d).max() < 1e-5
    assert P.e.react_now() == 0            # no reaction defined: a pass is a no-op
    P.close()


def test_particles_on_and_outside_the_box_boundary():
    m = util.melt(9, seed=2, trimers=False)
    L = m["box"][0]
    pos = m["pos"].copy()
    pos[0] = (0.0, 0.0, 0.0); pos[1] = (L, 0.5 * L, L); pos[2] = (-0.25 * L, 2.75 * L, 1.5 * L); pos[3] = (L * (1 - 1e-12), 1e-13, L)
    P = util.Pair(pos, m["box"], m["type"])
    st = P.e.get_particles(fields=("pos", "image"))
    assert (st["pos"] >= 0).all() and (st["pos"] < L).all()
    assert np.allclose(st["pos"] + st["image"] * m["box"], pos, atol=1e-7)
    r, e, f = util.lj_table()
    tab = P.add_table(r, e, f, 1)
    P.nb_tab([(0, 0)], tab, 2.5)
    a, b = P.e.pairs(), P.o.pairs()
    assert len(a) == len(b) and (a == b).all()
    P.close()


def test_all_pair_kernel_variants_agree():
    m = util.melt(14, seed=6)
    P = util.Pair(m["pos"], m["box"], m["type"], state=np.ones(len(m["pos"]), np.int32), resid=m["resid"])
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    r, e, f = util.lj_table()
    tab = P.add_table(r, e, f, 1)
    P.nb_tab(util.type_pairs(2), tab, 2.5)
    fe, fo = _forces(P)
    assert P.e.get_option("pair_kernel") == 3      # windowed-table kernel (round 2) is the default for all-tabulated cubic systems
    assert util.rel_force_err(fe, fo) < 1e-6
    # no thermostat -> the table window starts at row 0 and every listed pair inside the cutoff finds its row in shared memory:
    # the same arithmetic in the same order as the round-1 kernel (pair_kernel=2) -> bit-identical forces for every layout
    for opts in (dict(pair_nv=1), dict(pair_nv=3), dict(pair_nv=0, pair_ni=2), dict(pair_ni=4, pair_rep=8), dict(pair_rep=1),
                 dict(pair_kernel=2), dict(pair_kernel=2, pair_ni=2), dict(pair_kernel=2, pair_ni=4, tables_in_smem=0), dict(tables_in_smem=1)):
        for k, v in opts.items():
            P.e.set_option(k, v)
        P.e.compute_forces()
        f2 = P.e.get_particles(fields=("force",))["force"]
        assert (f2 == fe).all(), opts
    # rows read through the global-memory path (no shared-memory budget at all; a thermal window): same rows, the out-of-window
    # pairs are summed after the others -> equal within rounding, and still within the force bar
    P.e.set_option("pair_kernel", 3)
    for opts in (dict(pair_table_kb=0), dict(pair_table_kb=4), dict(pair_table_kb=-1)):
        for k, v in opts.items():
            P.e.set_option(k, v)
        P.e.compute_forces()
        f2 = P.e.get_particles(fields=("force",))["force"]
        assert util.rel_force_err(f2, fo) < 1e-6 and np.abs(f2 - fe).max() <= 1e-9 * np.abs(fe).max(), opts
    P.both("set_langevin", 1, 1.0, 1.0)        # thermostat on -> thermal window (U - Umin < 30 kT): the wall rows come from global memory
    P.e.compute_forces()
    f2 = P.e.get_particles(fields=("force",))["force"]
    assert util.rel_force_err(f2, fo) < 1e-6 and P.e.get_option("pair_table_rows") < 1500
    P.e.set_option("pair_kernel", 1)           # first-generation kernels: different arithmetic, same tolerance
    for bf in (1, 0):
        P.e.set_option("pair_branchfree", bf)
        P.e.compute_forces()
        f3 = P.e.get_particles(fields=("force",))["force"]
        assert util.rel_force_err(f3, fo) < 1e-6
    P.close()


def test_several_tables_and_a_type_pair_without_potential():
    m = util.melt(12, seed=8)
    n = len(m["pos"])
    rng = np.random.default_rng(3)
    types = m["type"].copy()
    types[rng.random(n) < 0.3] = 2
    types[rng.random(n) < 0.1] = 3
    P = util.Pair(m["pos"], m["box"], types)
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    r, e, f = util.lj_table()
    t1 = P.add_table(r, e, f, 1)
    t2 = P.add_table(r, 0.5 * e, 0.5 * f, 1)
    t3 = P.add_table(r, 1.7 * e, 1.7 * f, 1)
    a = P.e.add_nonbonded("Tabulated"); b = P.o.add_nonbonded(1)
    for (x, y), t, rc in (((0, 0), t1, 2.5), ((0, 1), t2, 2.5), ((1, 1), t3, 2.2), ((0, 2), t2, 1.9), ((2, 2), t1, 2.5), ((1, 3), t3, 2.5)):
        P.e.nb_set_tabulated(a, x, y, t, rc); P.o.nb_set_tab(b, x, y, t, rc)
    # pairs (1,2), (0,3), (2,3), (3,3) carry no potential at all
    fe, fo = _forces(P)
    assert P.e.get_option("pair_kernel") == 3 and P.e.get_option("pair_tables_resident") == 3
    assert util.rel_force_err(fe, fo) < 1e-6
    assert abs(P.e.energy(a) - P.o.energy(b)) <= 1e-8 * abs(P.o.energy(b))
    # only the hottest table fits (24 KB budget), nothing fits, round-1 kernel: same forces
    for opts in (dict(pair_table_kb=30), dict(pair_table_kb=0), dict(pair_kernel=2)):
        for k, v in opts.items():
            P.e.set_option(k, v)
        P.e.compute_forces()
        f2 = P.e.get_particles(fields=("force",))["force"]
        assert util.rel_force_err(f2, fo) < 1e-6 and np.abs(f2 - fe).max() <= 1e-9 * np.abs(fe).max(), opts
    P.close()


def test_reaction_without_candidates_and_inactive_reaction():
    m = util.melt(12, seed=5)
    n = len(m["pos"])
    P = util.Pair(m["pos"], m["box"], m["type"], state=np.ones(n, np.int32), resid=m["resid"])
    P.exclusions(util.exclusions_from(m["bonds"], m["angles"]))
    r, e, f = util.lj_table()
    tab = P.add_table(r, e, f, 1)
    P.nb_tab(util.type_pairs(2), tab, 2.5)
    rl = P.add_list(2, np.zeros((0, 2), np.int64))
    P.both("set_dt", 0.004)
    P.both("reaction_general", 1, 10, 1, 0)
    # cutoff below every pair distance -> no candidate; an inactive reaction -> no candidate either
    ra = P.e.add_reaction(0, 0, 1, 1, 1, 2, 1, 2, 1e6, 0.05, rl); rb = P.o.add_reaction(0, 0, 1, 1, 1, 2, 1, 2, 1e6, 0.05, rl)
    r2a = P.e.add_reaction(0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.2, rl, active=0); r2b = P.o.add_reaction(0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.2, rl, active=0)
    assert ra == rb and r2a == r2b
    assert P.e.react_now() == P.o.react() == 0
    assert P.e.list_size(rl) == 0
    P.e.reaction_set_active(r2a, 1); P.o.reaction_set_active(r2b, 1)
    na, nb = P.e.react_now(), P.o.react()
    assert na == nb > 0
    la, lb = P.e.list_get(rl, 2), P.o.list_get(rl, 2)
    assert (la[np.lexsort((la[:, 1], la[:, 0]))] == lb[np.lexsort((lb[:, 1], lb[:, 0]))]).all()
    P.close()
