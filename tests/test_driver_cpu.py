"""Host-side driver layer (no GPU needed): the reference's own two unit tests re-expressed, the shipped exclusion
list of examples/atrp_lj as a known answer for exclusion generation, type-id assignment from the shipped run log,
and the espressopp surface's constructibility rules."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def test_parse_exchange_reaction():
    # src/tests/test_reaction_parser.py:29-51
    from chemlab_b200.chemlab import reaction_parser as rp
    r, kind = rp.parse_exchange_equation("C(0,1):E(0,1) + W(0,1) -> A(1):Z(1) + E(1)")
    assert kind == rp.REACTION_EXCHANGE
    assert (r["type_1"]["name"], r["type_1"]["new_type"], r["type_1"]["min"], r["type_1"]["max"], r["type_1"]["delta"]) == ("C", "A", "0", "1", "1")
    assert (r["type_2"]["name"], r["type_2"]["new_type"], r["type_2"]["min"], r["type_2"]["max"], r["type_2"]["delta"]) == ("E", "E", "0", "1", "1")
    assert (r["type_3"]["name"], r["type_3"]["new_type"], r["type_3"]["min"], r["type_3"]["max"], r["type_3"]["delta"]) == ("W", "Z", "0", "1", "1")


def test_replicated_molecules():
    # src/tests/test_topology_reader.py:34-69 on the reference's own fixture (topol.top + diol_cg.itp + ter_cg.itp)
    from chemlab_b200.chemlab.gromacs_topology import GromacsTopology
    gt = GromacsTopology(os.path.join(GOLD, "parser", "topol.top"), generate_exclusions=True).read()
    for name, store in (("atoms", gt.atoms), ("bonds", gt.bonds), ("angles", gt.angles), ("dihedrals", gt.dihedrals), ("pairs", gt.pairs)):
        expected = sum(n * len(gt.gt.molecules_data[mol].get(name, [])) for mol, n in gt.gt.molecules)
        assert len(store) == expected, name
    assert len(gt.atoms) > 0 and len(gt.bonds) > 0


def test_atrp_lj_topology_and_exclusions_match_shipped_artefacts():
    from chemlab_b200.chemlab.gromacs_topology import GromacsTopology, gen_particle_list
    from chemlab_b200.chemlab.files_io import GROFile
    d = os.path.join(GOLD, "atrp_lj")
    gt = GromacsTopology(os.path.join(d, "topol.top")).read()
    # counts logged by the reference run (examples/atrp_lj/single:37,43-45): 6000 particles, 4000 bonds, 2000 angles
    assert (len(gt.atoms), len(gt.bonds), len(gt.angles)) == (6000, 4000, 2000)
    # exclusion_topol.list is what src/start_simulation.py:182-187 wrote from gt.exclusions
    want = sorted(tuple(int(x) for x in l.split()) for l in open(os.path.join(d, "exclusion_topol.list")) if l.strip())
    assert sorted(gt.exclusions) == want
    # type ids: MA ML first (molecule order), then the remaining [atomtypes] (run log :199-205: MA0 ML1 DA2 FA3 PA4 RA5 PL6)
    ids = gt.atomsym_atomtype
    assert ids["MA"] == 0 and ids["ML"] == 1 and set(ids) == {"MA", "ML", "DA", "FA", "PA", "RA", "PL"} and sorted(ids.values()) == list(range(7))
    conf = GROFile(os.path.join(d, "conf.gro")); conf.read()
    assert len(conf.atoms) == 6000 and abs(conf.box[0] - 28.11442) < 1e-5
    props, rows = gen_particle_list(conf, gt)
    assert props[:3] == ["id", "type", "pos"] and len(rows) == 6000 and rows[0][0] == 1
    # [atomstate] gives the initial chemical state per type (files_io.py:682-687)
    assert {r[6] for r in rows if r[1] == ids["MA"]} == {gt.gt.atomtypes["MA"].get("state", 0)}


def test_arg_file_convention(tmp_path):
    from chemlab_b200.chemlab import app_args
    f = tmp_path / "params"
    f.write_text("conf=conf.gro\ntop=topol.top\n# comment\nrun=2000 ; trailing\nskin=0.4\nkb=1.0\nreactions=atrp.cfg\n")
    a = app_args._args().parse_args(["@%s" % f])
    assert a.run == 2000 and float(a.skin) == 0.4 and a.kb == 1.0 and a.reactions == "atrp.cfg" and a.thermostat == "lv"
    shipped = app_args._args().parse_args(["@%s" % os.path.join(GOLD, "atrp_lj", "params")])
    assert shipped.lj_cutoff == 2.5 and shipped.dt == 0.0025 and shipped.maximum_conversion == "PL(1):1200:2000"


def test_reaction_config_of_atrp_lj():
    from chemlab_b200.chemlab import reaction_parser as rp
    c = rp.parse_config(os.path.join(GOLD, "atrp_lj", "atrp.cfg"))
    assert c["general"]["interval"] == 200 and c["general"]["nearest"] is True      # bool('0') quirk preserved
    g = c["reactions"]["reaction_1"]
    assert g["potential"] == "Harmonic" and g["potential_options"] == {"K": "30.0", "r0": "0.97"}
    assert [r["equation"].split()[0] for r in g["reaction_list"]] == ["FA(3,", "DA(3,", "FA(3,", "DA(3,"]
    assert g["extensions"]["atrp"]["class"] == "ATRPActivator" and g["extensions"]["change_neighbour_type"]["class"] == "ChangeNeighboursProperty"


def test_out_of_scope_names_are_constructible_but_refuse_to_run():
    import chemlab_b200.espressopp as es
    s = es.System()
    s.rng = es.esutil.RNG(1); s.bc = es.bc.OrthorhombicBC(s.rng, (10, 10, 10)); s.skin = 0.3
    s.storage = es.storage.DomainDecomposition(s, (1, 1, 1), (3, 3, 3))
    integ = es.integrator.VelocityVerlet(s)
    vl = es.VerletList(s, cutoff=2.5, exclusionlist=es.DynamicExcludeList(integ, []))
    capped = es.interaction.VerletListTabulatedCapped(vl)            # gromacs_topology.py:513 builds it unconditionally
    with pytest.raises(NotImplementedError):
        capped.setPotential(type1=0, type2=0, potential=None)
    with pytest.raises(NotImplementedError):
        es.io.DumpH5MD(s, integ).dump()


def test_molecule_exclusions_nrexcl():
    from chemlab_b200.chemlab.gromacs_topology import GromacsTopology as G
    chain = [(1, 2), (2, 3), (3, 4), (4, 5)]
    assert G.molecule_exclusions(chain, 1) == {(1, 2), (2, 3), (3, 4), (4, 5)}
    assert G.molecule_exclusions(chain, 2) == {(1, 2), (2, 3), (3, 4), (4, 5), (1, 3), (2, 4), (3, 5)}
    assert (1, 4) in G.molecule_exclusions(chain, 3) and (1, 5) not in G.molecule_exclusions(chain, 3)


def test_driver_host_logic_on_the_oracle_backend(tmp_path, monkeypatch):
    """The whole driver (arg-file, topology, exclusions, reactions, ATRP activator, hooks, observers, outputs) with the engine
    swapped for the CPU checker (oracle/engine_adapter.py): host logic only, no GPU.  The products of the reference driver must
    appear (src/start_simulation.py:800-1081) and must be readable by our own readers."""
    import shutil
    import sys
    sys.path.insert(0, HERE)
    import chemlab_b200.espressopp._context as C
    from oracle.engine_adapter import OracleEngine
    from chemlab_b200 import start_simulation as S
    from chemlab_b200.chemlab.files_io import GROFile
    from chemlab_b200.chemlab.gromacs_topology import GromacsTopology
    d = str(tmp_path / "atrp")
    shutil.copytree(os.path.join(GOLD, "atrp_lj"), d)
    monkeypatch.chdir(d)
    monkeypatch.setattr(C, "Engine", OracleEngine)
    r = S.main(["@params", "--rng_seed", "42", "--run", "600", "--start_ar", "200", "--energy_collect", "200", "--save_before_reaction", "True"])
    assert r["steps"] == 600
    out = set(os.listdir("data"))
    for name in ("cpc01_42_confout.gro", "cpc01_42_before_reaction_confout.gro", "cpc01_42_output_topol.top", "cpc01_42_state.dat",
                 "cpc01_42_bonds.dat", "cpc01_42_angles.dat", "cpc01_42_reaction_counters", "cpc01_42_intra_inter_counters", "cpc01_42_benchmark.csv",
                 "cpc01_42_atrp.cfg", "cpc01_42_whole_confout.gro",
                 "cpc01_energy_42.csv", "cpc01params.out", "cpc01_42_atrp_stats.dat", "cpc01_42_topology.dat", "cpc01_42_res_topology.dat",
                 "cpc01_42_residue_list.dat", "cpc01_42_benchmark.pck"):
        assert name in out, name
    # _bonds.dat / _angles.dat rows: ids, func, parameters, origin (src/start_simulation.py:897-988)
    brow = [l.split() for l in open(os.path.join("data", "cpc01_42_bonds.dat"))]
    assert len(brow) >= 4000 and brow[0] == ["1", "2", "1", "0.97", "60.0", ";", "dynamic"]
    assert all(r[-2:] == [";", "dynamic"] or r[-3:-1] == [";", "chem"] for r in brow) and not any("MISSING" in r for r in brow)
    arow = [l.split() for l in open(os.path.join("data", "cpc01_42_angles.dat"))]
    assert len(arow) >= 2000 and arow[0] == ["1", "2", "3", "1", "180.0", "2.5", ";", "dynamic"]
    import pickle
    bp = pickle.load(open(os.path.join("data", "cpc01_42_benchmark.pck"), "rb"))
    assert set(bp) == {"traj_timers", "topol_timers", "integrator_timers", "extension_timers", "verlet_list"}      # src/start_simulation.py:1062-1076
    topo = [l.split(":") for l in open(os.path.join("data", "cpc01_42_topology.dat"))]
    assert len(topo) == 6000 and all(len(nb.split()) in (1, 2, 3) for _, nb in topo)        # trimers + the bonds the reactions added
    resl = open(os.path.join("data", "cpc01_42_residue_list.dat")).read().splitlines()
    assert len(resl) == 2000 and all(len(l.split(":")[1].split()) == 3 for l in resl)
    stats = np.loadtxt(os.path.join("data", "cpc01_42_atrp_stats.dat"), ndmin=2)
    assert len(stats) >= 1 and (stats[:, 1] + stats[:, 2]).sum() > 0                     # the activator fired and changed some chain ends
    g = GROFile(os.path.join("data", "cpc01_42_confout.gro")); g.read()
    # the end configuration is the input file with new (folded) positions: same title, atom and residue names (:1008-1012)
    g0 = GROFile(os.path.join(d, "conf.gro")); g0.read()
    assert len(g.atoms) == 6000 and g.title == g0.title and all(g.atoms[k].name == g0.atoms[k].name and g.atoms[k].chain_name == g0.atoms[k].chain_name for k in g0.atoms)
    assert all(0.0 <= x < 28.11442 + 1e-3 for k in (1, 3000, 6000) for x in g.atoms[k].position)
    st = np.loadtxt(os.path.join("data", "cpc01_42_state.dat"))
    assert st.shape == (6000, 4) and len(set(st[:, 1].astype(int))) >= 4                                  # current types live in _state.dat
    csv = open(os.path.join("data", "cpc01_energy_42.csv")).read().splitlines()
    # columns and rows of the reference's SystemMonitor (src/start_simulation.py:446-569,705): T, Ekin, the interactions sorted by
    # label, the conversion observers (cr_<type id>_<state> of maximum_conversion=PL(1):...), count_<k> per reaction list; the row
    # of step 0 is dumped before the loop, then one row every min(reaction interval, energy_collect) steps
    assert csv[0].split("\t") == ["step", "time", "T", "Ekin", "chem_fpl_reaction_1", "dyn_angles_0", "dyn_bonds_0", "lj", "cr_6_1", "count_0"]
    assert [int(l.split("\t")[0]) for l in csv[1:]] == [0, 200, 400, 600]
    assert "cpc01_42_before_reaction_confout.gro" in out                               # written unconditionally (:742-746)
    assert open(os.path.join("data", "cpc01_42_benchmark.csv")).read().split()[:2] == ["1", "6000"]
    gt = GromacsTopology(os.path.join("data", "cpc01_42_output_topol.top")).read()
    assert len(gt.atoms) == 6000 and len(gt.bonds) >= 4000 and len(gt.angles) >= 2000
    # the output topology carries the force-field sections of the input (files_io.py:540-575) and, with the end configuration,
    # restarts the driver: the reaction bonds are now ordinary bonds, their generated angles ordinary angles, both excluded
    sections = [l.strip() for l in open(os.path.join("data", "cpc01_42_output_topol.top")) if l.startswith("[")]
    assert sections == ["[ defaults ]", "[ atomtypes ]", "[ bondtypes ]", "[ angletypes ]", "[ atomstate ]", "[ moleculetype ]",
                        "[ atoms ]", "[ bonds ]", "[ angles ]", "[ dihedrals ]", "[ pairs ]", "[ system ]", "[ molecules ]"]
    nb, na = len(gt.bonds), len(gt.angles)
    shutil.copy(os.path.join("data", "cpc01_42_output_topol.top"), "restart.top")
    shutil.copy(os.path.join("data", "cpc01_42_confout.gro"), "restart.gro")
    r2 = S.main(["@params", "--rng_seed", "43", "--run", "100", "--top", "restart.top", "--conf", "restart.gro", "--exclusion_list", "none.list",
                 "--reactions", "", "--energy_collect", "100", "--int_step", "100"])
    assert r2["steps"] == 100 and (len(r2["topology"].bonds), len(r2["topology"].angles)) == (nb, na)
    assert len(r2["system"]._ctx.engine.get_exclusions()) >= 6000 + (nb - 4000)


def test_coulomb_label_for_neutral_systems():
    """gromacs_topology.py:866-878: with coulomb_cutoff > 0 the reference registers a `coulomb` interaction for every type pair.
    All shipped coarse-grained beads carry q = 0, so the term is identically zero: the label and the energy column exist, a
    charged system is refused (Coulomb pair forces are outside north_star)."""
    from chemlab_b200 import espressopp
    from chemlab_b200.espressopp import interaction as I, analysis as A

    def build(q):
        system = espressopp.System()
        system.bc = espressopp.bc.OrthorhombicBC(system.rng, (10.0, 10.0, 10.0))
        system.skin = 0.3
        system.storage = espressopp.storage.DomainDecomposition(system, (1, 1, 1), (3, 3, 3))
        system.storage.addParticles([[1, 0, espressopp.Real3D(1, 1, 1), 1.0, q], [2, 0, espressopp.Real3D(2, 1, 1), 1.0, 0.0]], "id", "type", "pos", "mass", "q")
        vl = espressopp.VerletList(system, cutoff=2.5)
        c = I.VerletListCoulombTruncated(vl)
        c.setPotential(type1=0, type2=0, potential=I.CoulombTruncated(prefactor=138.935485, cutoff=0.9))
        system.addInteraction(c, "coulomb")
        return system, c
    system, c = build(0.0)
    c._attach(None)                                   # neutral: nothing to upload
    assert c.computeEnergy() == 0.0 and system.getNameOfInteraction(0) == "coulomb"
    assert A.PotentialEnergy(system, c)._inter is c
    system, c = build(0.5)
    with pytest.raises(NotImplementedError):
        c._attach(None)


def test_14_pairs_become_lj_pair_lists(tmp_path, monkeypatch):
    """[ pairs ] -> lj14_<k> / dyn_lj14_<k> interactions (gromacs_topology.py:1314-1411): parameters from the pair line or from
    the combination rule with fudgeLJ; pairs touching a reactive type go to the type-dispatched list."""
    from chemlab_b200 import espressopp
    from chemlab_b200.chemlab import gromacs_topology as G

    class Args:
        lj_cutoff = 1.2

    class GT:       # the few attributes set_pair_interactions reads
        pass
    gt = GT(); gt.gt = GT()
    gt.gt.defaults = {"combinationrule": 2, "fudgeLJ": 0.5, "gen-pairs": True}
    gt.gt.atomtypes = {"A": {"sigma": 0.3, "epsilon": 1.0}, "B": {"sigma": 0.5, "epsilon": 4.0}, "R": {"sigma": 0.4, "epsilon": 1.0}}
    gt.atomsym_atomtype = {"A": 0, "B": 1, "R": 2}
    gt.used_atomsym_atomtype = dict(gt.atomsym_atomtype)
    gt.atoms = {1: {"type_id": 0}, 2: {"type_id": 1}, 3: {"type_id": 0}, 4: {"type_id": 2}, 5: {"type_id": 1}}
    gt.pairs = {(1, 2): ["1"], (3, 5): ["1"], (1, 3): ["1", "0.35", "0.7"], (2, 4): ["1"]}
    system = espressopp.System()
    system.bc = espressopp.bc.OrthorhombicBC(system.rng, (10.0, 10.0, 10.0))
    system.storage = espressopp.storage.DomainDecomposition(system, (1, 1, 1), (3, 3, 3))
    dfpl, static = G.set_pair_interactions(system, gt, Args(), {2})
    names = [system.getNameOfInteraction(k) for k in range(system.getNumberOfInteractions())]
    assert names == ["lj14_0", "lj14_1", "dyn_lj14_2"]
    assert sorted(map(tuple, dfpl.getAllBonds())) == [(2, 4)] and len(static) == 2
    inters = system.getAllInteractions() if hasattr(system, "getAllInteractions") else None
    # explicit parameters are used as given; generated ones follow rule 2 (arithmetic sigma, geometric epsilon) times fudgeLJ
    pots = sorted((i._pot.sigma, i._pot.epsilon) for i, _ in system._ctx.interactions if getattr(i, "_pot", None) is not None)
    assert pots[0] == (0.35, 0.7) and abs(pots[1][0] - 0.4) < 1e-12 and abs(pots[1][1] - 0.5 * 2.0) < 1e-12
    gt.pairs = {}
    assert G.set_pair_interactions(system, gt, Args(), set()) == (None, [])


def run_dacron_restrict(tmp, backend, steps, seed="7"):
    """examples/dacron/restrict through the driver (shared by the CPU test below and tests/test_gpu_zz_restrict.py): the shipped
    connectivity map, topology, coordinates and arg-file; the reaction interval is shortened to 100 steps and the rate raised so
    that every allowed pair inside the cut-off reacts (the shipped p = 0.005 per pass is meant for 2e7-step runs)."""
    import sys
    sys.path.insert(0, HERE)
    from chemlab_b200 import synthetic
    import chemlab_b200.espressopp._context as C
    from chemlab_b200 import start_simulation as S
    d = synthetic.prepare_example(os.path.join(GOLD, "dacron_restrict"), os.path.join(tmp, "restrict_" + backend), "dacron_restrict")
    cwd = os.getcwd()
    os.chdir(d)
    real = C.Engine
    try:
        cfg = open("reaction.cfg").read()
        assert "connectivity_map:connections.list" in cfg and cfg.count("rate: 0.005") == 2
        open("reaction.cfg", "w").write(cfg.replace("interval: 1000", "interval: 100").replace("rate: 0.005", "rate: 100.0"))
        if backend == "oracle":
            from oracle.engine_adapter import OracleEngine
            C.Engine = OracleEngine
        r = S.main(["@params", "--run", str(steps), "--rng_seed", seed, "--t_hybrid_bond", "0", "--gen_velocity", "True",
                    "--int_step", "100", "--energy_collect", "100"])
        e = r["system"]._ctx.engine
        g = e.get_particles(fields=("type", "state", "mass"))
        bonds = np.concatenate([np.asarray(f.fpl.getAllBonds(), np.int64).reshape(-1, 2) for f in r["chem_fpls"]])
        conn = {tuple(sorted(int(x) for x in l.split())) for l in open("connections.list") if l.strip()}
        names = [r["system"].getNameOfInteraction(k) for k in range(r["system"].getNumberOfInteractions())]
        return dict(g=g, bonds=bonds, conn=conn, steps=r["steps"], names=names, reactions=r.get("reactions"))
    finally:
        C.Engine = real
        os.chdir(cwd)


def test_dacron_restrict_driver_on_the_oracle_backend(tmp_path):
    """RestrictReaction end to end (reaction_setup.py:74-75,115-126; examples/dacron/restrict): every bond the reactions form is
    a line of connections.list, and the run forms at least one (1998 allowed pairs among 4000 beads: about one pair per pass
    comes within the 0.48 nm cut-off)."""
    a = run_dacron_restrict(str(tmp_path), "oracle", 400)
    assert a["steps"] == 400 and len(a["conn"]) == 1998
    assert len(a["bonds"]) >= 1
    assert all(tuple(sorted(b)) in a["conn"] for b in a["bonds"].tolist())
    t = a["g"]["type"]
    assert (t >= 3).sum() >= 1          # products C / E exist


def run_mf(tmp, backend, steps, seed="3", interval=None):
    """examples/mf/espp_cg_1 as shipped through the driver (shared with tests/test_gpu_zz_restrict.py, which shortens the
    reaction interval so that several passes fall inside the time over which two fp64 trajectories stay together)."""
    import sys
    sys.path.insert(0, HERE)
    from chemlab_b200 import synthetic
    import chemlab_b200.espressopp._context as C
    from chemlab_b200 import start_simulation as S
    d = synthetic.prepare_example(os.path.join(GOLD, "mf"), os.path.join(tmp, "mf_" + backend), "mf")
    cwd = os.getcwd()
    os.chdir(d)
    real = C.Engine
    try:
        if backend == "oracle":
            from oracle.engine_adapter import OracleEngine
            C.Engine = OracleEngine
        extra = []
        if interval:
            cfg = open("reaction.cfg").read()
            assert "interval: 1000" in cfg
            open("reaction.cfg", "w").write(cfg.replace("interval: 1000", "interval: %d" % interval))
            extra = ["--int_step", str(interval)]
        r = S.main(["@params", "--run", str(steps), "--rng_seed", seed, "--start_ar", "0"] + extra)
        g = r["system"]._ctx.engine.get_particles(fields=("type", "state", "mass"))
        bonds = np.concatenate([np.asarray(f.fpl.getAllBonds(), np.int64).reshape(-1, 2) for f in r["chem_fpls"]])
        return dict(g=g, bonds=bonds, steps=r["steps"])
    finally:
        C.Engine = real
        os.chdir(cwd)


def test_mf_driver_on_the_oracle_backend(tmp_path):
    """examples/mf/espp_cg_1 as shipped: 1000 single-bead molecules, A(0,3) + A(0,3) -> A(1):A(1), cutoff 0.5, intramolecular: 0,
    interval 1000 (reaction.cfg).  Size-independent properties of the result: a bead's state counts its new bonds and stays below
    3, and -- because a bond inside one molecule is forbidden and molecule ids merge with every bond -- the bond graph is a forest."""
    a = run_mf(str(tmp_path), "oracle", 3000)
    b = a["bonds"]
    assert a["steps"] == 3000 and len(b) > 20
    deg = np.bincount(b.ravel(), minlength=1001)
    st = a["g"]["state"]
    assert deg.max() <= 3 and (deg[1:] == st).all()
    parent = list(range(1001))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]; x = parent[x]
        return x
    for i, j in b.tolist():
        ri, rj = find(i), find(j)
        assert ri != rj, "a reaction closed a ring inside one molecule (%d, %d)" % (i, j)
        parent[ri] = rj


def run_pccg_lj(tmp, backend, steps, seed="3"):
    """examples/pccg_lj/chemical_reactions as shipped (hooks.py re-authored for Python 3) through the driver."""
    import shutil
    import sys
    sys.path.insert(0, HERE)
    import chemlab_b200.espressopp._context as C
    from chemlab_b200 import start_simulation as S
    d = os.path.join(tmp, "pccg_" + backend)
    shutil.copytree(os.path.join(GOLD, "pccg_lj"), d)
    cwd = os.getcwd()
    os.chdir(d)
    real = C.Engine
    try:
        if backend == "oracle":
            from oracle.engine_adapter import OracleEngine
            C.Engine = OracleEngine
        r = S.main(["@params", "--run", str(steps), "--rng_seed", seed, "--energy_collect", "200"])
        system = r["system"]
        g = system._ctx.engine.get_particles(fields=("pos", "type", "state", "mass"))
        bonds = np.concatenate([np.asarray(f.fpl.getAllBonds(), np.int64).reshape(-1, 2) for f in r["chem_fpls"]])
        names = [system.getNameOfInteraction(k) for k in range(system.getNumberOfInteractions())]
        return dict(g=g, bonds=bonds, steps=r["steps"], names=names, dir=d, system=system, files=sorted(os.listdir(os.path.join(d, "data"))))
    finally:
        C.Engine = real
        os.chdir(cwd)


def test_pccg_lj_driver_on_the_oracle_backend(tmp_path):
    """examples/pccg_lj/chemical_reactions: pair-specific Lennard-Jones, FENE + LJ bonds for the monomers and for the reaction bonds
    (group potential FENELennardJones), Cosine angles generated by the TopologyManager, ATRPActivator, all five user hooks of the
    reference driver (src/start_simulation.py:215-228) incl. `import espressopp` inside hooks.py and analysis.AngleDistribution."""
    import chemlab_b200.espressopp as es
    a = run_pccg_lj(str(tmp_path), "oracle", 600)
    assert a["steps"] == 600 and a["names"][:3] == ["chem_fpl_reaction_1", "coulomb", "lj"] and "dyn_angles_0" in a["names"]
    assert len(a["bonds"]) >= 3
    assert any(f.endswith("_atrp_stats.dat") for f in a["files"])
    hist = np.loadtxt(os.path.join(a["dir"], "output_angle.csv"))           # hook_before_sim + hook_at_step + hook_end ran
    assert hist.shape == (100, 2) and hist[:, 1].sum() > 0
    # AngleDistribution against a direct loop over the bond graph (monomer bonds + reaction bonds) at the final positions
    system = a["system"]
    obs = es.analysis.AngleDistribution(system); obs.load_from_topology_manager(system.topology_manager)
    got = np.array(obs.compute(100))
    tm_bonds = np.concatenate([np.asarray(f.getAllBonds(), np.int64).reshape(-1, 2) for f in system.topology_manager._observed])
    assert len(tm_bonds) == 2000 + len(a["bonds"])
    adj = {}
    for i, j in tm_bonds.tolist():
        adj.setdefault(i, []).append(j); adj.setdefault(j, []).append(i)
    pos, box = a["g"]["pos"], 26.150192
    want = np.zeros(100, np.int64)
    for j, nb in adj.items():
        for x in range(len(nb)):
            for y in range(x + 1, len(nb)):
                u = pos[nb[x] - 1] - pos[j - 1]; v = pos[nb[y] - 1] - pos[j - 1]
                u -= box * np.rint(u / box); v -= box * np.rint(v / box)
                th = np.arccos(np.clip(u @ v / np.sqrt((u @ u) * (v @ v)), -1, 1))
                want[min(int(th / (np.pi / 100)), 99)] += 1
    assert want.sum() > 0 and (got == want).all()


def test_maximum_conversion_stops_the_run_like_the_reference(tmp_path, monkeypatch):
    """src/start_simulation.py:757-770 + src/tools.py:102-180: the run ends as soon as ANY stop criterion is reached (checked at the
    top of an outer iteration once the reactions are on); `A+B(1)` criteria are ONE observable whose column is cr_<A>_<B>; with
    eq_steps the loop goes on for int(eq_steps / number of outer iterations) more iterations with the reactions disconnected."""
    import shutil
    import sys
    sys.path.insert(0, HERE)
    import chemlab_b200.espressopp._context as C
    from oracle.engine_adapter import OracleEngine
    from chemlab_b200 import start_simulation as S
    monkeypatch.setattr(C, "Engine", OracleEngine)
    res = {}
    for tag, extra in (("stop", []), ("eq", ["--eq_steps", "16"])):
        d = str(tmp_path / tag)
        shutil.copytree(os.path.join(GOLD, "atrp_lj"), d)
        monkeypatch.chdir(d)
        # 3 PL beads are in state 1 after the pass at step 600 (see the test above): the first criterion can never be met, the second is
        r = S.main(["@params", "--rng_seed", "42", "--run", "1600", "--start_ar", "200", "--energy_collect", "200",
                    "--maximum_conversion", "MA:5000:6000,PL(1)+FA(7):2:2000"] + extra)
        hdr = open(os.path.join("data", "cpc01_energy_42.csv")).readline().split()
        assert "cr_0" in hdr and "cr_PL_FA" in hdr
        bonds = np.concatenate([np.asarray(f.fpl.getAllBonds(), np.int64).reshape(-1, 2) for f in r["chem_fpls"]])
        res[tag] = (r["steps"], len(bonds))
    assert res["stop"] == (600, 3)
    # 8 outer iterations -> eq_run = int(16 / 8) = 2 more iterations, without reactions: the bond count stays
    assert res["eq"] == (1000, 3)


def test_python2_dict_order_of_the_reference_is_reproduced():
    """Where the Python-2 reference iterates a plain dict, the order shows in its results.  Known answer: the type ids of the
    shipped run log examples/atrp_lj/single:199-205 (MA0 ML1 DA2 FA3 PA4 RA5 PL6) for the [ atomtypes ] file order MA ML PA FA DA RA PL
    (examples/atrp_activator/topol.top:3-11 without the later row I).  The same rule orders the reaction groups."""
    from chemlab_b200.chemlab.py2compat import py2_dict_order
    from chemlab_b200.chemlab import reaction_parser as rp
    assert py2_dict_order(["MA", "ML", "PA", "FA", "DA", "RA", "PL"]) == ["MA", "ML", "DA", "FA", "PA", "RA", "PL"]
    assert py2_dict_order(["a", "a", "b"]) in (["a", "b"], ["b", "a"])                 # re-insertion keeps one entry
    big = ["k%d" % i for i in range(200)]
    assert sorted(py2_dict_order(big)) == sorted(big)                                 # several resizes, nothing lost
    c = rp.parse_config(os.path.join(GOLD, "rim135", "reaction.cfg"))
    assert list(c["reactions"]) == ["reaction_2", "reaction_1"]


def test_driver_calls_bind_to_the_cuda_engine_signatures(tmp_path, monkeypatch):
    """The CPU suite drives the oracle adapter; on the GPU box the same driver drives chemlab_b200.Engine.  Every call the driver
    makes here (method name, positional and keyword arguments) must bind to the signature of the ctypes mirror of the C-ABI, so a
    driver path that was only ever exercised on the oracle cannot fail on the GPU for a missing method or argument."""
    import inspect
    import oracle.engine_adapter as EA
    from chemlab_b200.engine import Engine
    calls = {}
    real = EA.OracleEngine

    class Proxy:
        def __init__(self, *a, **k):
            inspect.signature(Engine.__init__).bind(None, *a, **k)
            object.__setattr__(self, "_o", real(*a, **k))

        def __getattr__(self, name):
            attr = getattr(self._o, name)
            if not callable(attr):
                return attr

            def f(*a, **k):
                calls.setdefault(name, []).append((a, k))
                return attr(*a, **k)
            return f

        def __setattr__(self, k, v):
            setattr(self._o, k, v)
    monkeypatch.setattr(EA, "OracleEngine", Proxy)
    run_pccg_lj(str(tmp_path), "oracle", 200)
    run_dacron_restrict(str(tmp_path), "oracle", 100)
    assert len(calls) >= 35 and {"reaction_define_connections", "atrp_now", "set_cap_force", "get_particles", "energy"} <= set(calls)
    for name, lst in calls.items():
        m = getattr(Engine, name, None)
        assert m is not None, "Engine has no method %s" % name
        for a, k in lst[:40]:
            inspect.signature(m).bind(None, *a, **k)


def test_potentials_equal_the_formulas_of_the_reference_documentation():
    """doc/topology.rst:7-129 states every analytic potential as a function of the parameters written in the topology file.  The
    chain under test is the one a run uses: topology parameters -> this driver's conversion (`_bond_pot`, `_angle_pot`,
    `_dihedral_pot`: K halved where chemlab halves it, degrees -> radians) -> the espressopp surface's parameter hand-down -> the
    oracle's evaluation (which the CUDA kernels are tested against).  Cosine follows chemlab's CODE, which does not halve K although
    the documentation says so (REFERENCE_UNVERIFIED U29)."""
    import math
    from oracle import pyoracle
    from chemlab_b200.engine import POT
    from chemlab_b200.chemlab import gromacs_topology as G

    def energy(arity, pot, pos):
        n = len(pos)
        o = pyoracle.Oracle(n, [30.0, 30.0, 30.0], 2.5, 0.3, seed=1)
        o.set_particles(np.asarray(pos, float) + 10.0, None, np.ones(n), None, np.zeros(n, np.int32), None, None)
        lst = o.add_list(arity); o.list_add(lst, [list(range(n))])
        h = o.add_bonded(lst, 0)
        o.bonded_set_potential(h, (), POT[pot.kind], pot.params())
        o.compute_forces()
        return o.energy(h)

    def angle_pos(theta):
        return [[1.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.9 * math.cos(theta), 0.9 * math.sin(theta), 0.0]]

    def dihedral_pos(phi):
        return [[1.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, math.cos(phi), math.sin(phi)]]
    close = lambda a, b: abs(a - b) <= 1e-11 * max(1.0, abs(b))
    for r in (0.8, 0.97, 1.3):
        r0, K = 0.97, 60.0
        assert close(energy(2, G._bond_pot(1, [str(r0), str(K)]), [[0, 0, 0], [r, 0, 0]]), 0.5 * K * (r - r0) ** 2)                    # eq1
        b, K = 1.5, 30.0
        fene = -0.5 * K * b * b * math.log(1.0 - r * r / (b * b))
        assert close(energy(2, G._bond_pot(7, [str(b), str(K)]), [[0, 0, 0], [r, 0, 0]]), fene)                                         # eqFENE
        sig, eps = 1.0, 1.2
        lj = 4 * eps * ((sig / r) ** 12 - (sig / r) ** 6)
        assert close(energy(2, G._bond_pot(9, [str(b), str(K), str(sig), str(eps)]), [[0, 0, 0], [r, 0, 0]]), fene + lj)                # eqFENELJ
    for deg in (95.0, 120.0, 170.0):
        th, th0, K = math.radians(deg), math.radians(110.0), 80.0
        assert close(energy(3, G._angle_pot(1, ["110.0", str(K)]), angle_pos(th)), 0.5 * K * (th - th0) ** 2)                          # eq2
        assert close(energy(3, G._angle_pot(11, ["110.0", str(K)]), angle_pos(th)), K * (1.0 + math.cos(th - th0)))                    # eq3 without its 1/2: U29
    phi_of = None
    for deg in (-150.0, -20.0, 75.0):
        phi0, K = math.radians(30.0), 12.0
        e = energy(4, G._dihedral_pot(12, ["30.0", str(K)]), dihedral_pos(math.radians(deg)))
        # the sign convention of phi (U15) is not part of this test: the quadruple is built for +deg or -deg
        assert any(close(e, 0.5 * K * (math.radians(s * deg) - phi0) ** 2) for s in (1.0, -1.0)), (deg, e)                             # eq7


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference examples are not mounted")
def test_shipped_python2_hook_file_runs_unmodified(tmp_path, monkeypatch, capsys):
    """examples/pccg_lj/chemical_reactions/hooks.py exactly as shipped (Python 2: a print statement; `import espressopp`; module-level
    state shared by hook_before_sim / hook_at_step / hook_end through `global`): the driver rewrites the print statement, and the hook
    file drives analysis.AngleDistribution and writes output_angle.csv."""
    import shutil
    import sys
    sys.path.insert(0, HERE)
    import chemlab_b200.espressopp._context as C
    from oracle.engine_adapter import OracleEngine
    from chemlab_b200 import start_simulation as S
    d = str(tmp_path / "pccg")
    shutil.copytree(os.path.join(GOLD, "pccg_lj"), d)
    shutil.copy("/root/reference/examples/pccg_lj/chemical_reactions/hooks.py", os.path.join(d, "hooks.py"))
    assert "print res_ids" in open(os.path.join(d, "hooks.py")).read()
    monkeypatch.chdir(d)
    monkeypatch.setattr(C, "Engine", OracleEngine)
    r = S.main(["@params", "--rng_seed", "3", "--run", "400", "--energy_collect", "200"])
    out = capsys.readouterr().out
    assert r["steps"] == 400 and "Python-2 print statements" in out and "Activated 20 monomers" in out
    hist = np.loadtxt("output_angle.csv")
    assert hist.shape == (100, 2)
