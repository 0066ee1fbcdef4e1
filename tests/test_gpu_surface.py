"""The espressopp-style surface on the GPU engine vs the same surface on the oracle, with tables SHIPPED by the
reference (tests/golden/*.pot = examples/dacron/.../table_A_A.pot, examples/hyperbranched/table_b0.pot): file-based
Tabulated potentials (itype 1 and Akima itype 2), FixedPairList interactions, exclusions, Langevin run, observables."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _build(backend):
    import chemlab_b200.espressopp as es
    import chemlab_b200.espressopp._context as C
    real = C.Engine
    if backend == "oracle":
        from oracle.engine_adapter import OracleEngine
        C.Engine = OracleEngine
    try:
        rng = np.random.default_rng(3)
        ns, a = 14, 0.45                                       # dacron-like number density (~11 beads / nm^3), nm units
        L = ns * a
        g = np.arange(ns)
        z, y, x = np.meshgrid(g, g, g, indexing="ij")
        pos = (np.stack([x.ravel(), y.ravel(), z.ravel()], 1) + 0.5) * a + rng.uniform(-0.04, 0.04, (ns ** 3, 3))
        n = len(pos)
        ids = np.arange(1, n + 1)                              # .gro numbering starts at 1
        first = ids[(x.ravel() % 2 == 0)]
        bonds = [(int(i), int(i) + 1) for i in first]          # dimers along x
        vel = rng.normal(0, 1.0, (n, 3))
        s = es.System()
        s.rng = es.esutil.RNG(7)
        s.bc = es.bc.OrthorhombicBC(s.rng, (L, L, L))
        s.skin = 0.1
        s.storage = es.storage.DomainDecomposition(s, (1, 1, 1), (4, 4, 4))
        integ = es.integrator.VelocityVerlet(s)
        integ.dt = 0.001
        rows = [(int(ids[k]), 0, es.Real3D(*pos[k]), 50.0, 0.0, int((ids[k] - 1) // 2), 0, es.Real3D(*vel[k])) for k in range(n)]
        s.storage.addParticles(rows, "id", "type", "pos", "mass", "q", "res_id", "state", "v")
        s.storage.decompose()
        excl = es.DynamicExcludeList(integ, bonds)
        vl = es.VerletList(s, cutoff=1.4, exclusionlist=excl)
        nb = es.interaction.VerletListTabulated(vl)
        nb.setPotential(type1=0, type2=0, potential=es.interaction.Tabulated(itype=1, filename=os.path.join(GOLD, "table_A_A.pot"), cutoff=1.4))
        s.addInteraction(nb, "lj-tab")
        half = len(bonds) // 2
        f1, f2 = es.FixedPairList(s.storage), es.FixedPairList(s.storage)
        f1.addBonds(bonds[:half]); f2.addBonds(bonds[half:])
        b1 = es.interaction.FixedPairListTabulated(s, f1, es.interaction.Tabulated(itype=1, filename=os.path.join(GOLD, "table_b0.pot")))
        b2 = es.interaction.FixedPairListTabulated(s, f2, es.interaction.Tabulated(itype=2, filename=os.path.join(GOLD, "table_b0.pot")))
        s.addInteraction(b1, "bonds_lin"); s.addInteraction(b2, "bonds_akima")
        th = es.integrator.LangevinThermostat(s)
        th.temperature = 2.5; th.gamma = 5.0
        integ.addExtension(th)
        e_before = [es.analysis.PotentialEnergy(s, s.getInteraction(k)).compute() for k in range(3)]
        eng = s._ctx.engine
        eng.compute_forces()
        f_before = eng.get_particles(fields=("force",))["force"].copy()
        integ.run(60)
        out = eng.get_particles(fields=("pos", "vel"))
        T = es.analysis.Temperature(s).compute()
        return dict(e=e_before, f=f_before, pos=out["pos"], T=T, L=L, nb=f1.totalSize() + f2.totalSize(), step=integ.step)
    finally:
        C.Engine = real


def test_surface_with_shipped_tables_matches_oracle():
    import clb_testutil as util
    a = _build("gpu")
    b = _build("oracle")
    assert a["step"] == b["step"] == 60 and a["nb"] == b["nb"] > 1000
    for ea, eb in zip(a["e"], b["e"]):
        assert abs(ea - eb) <= 1e-8 * abs(eb), (ea, eb)
    assert util.rel_force_err(a["f"], b["f"]) < 1e-6
    d = a["pos"] - b["pos"]
    d -= a["L"] * np.rint(d / a["L"])
    assert np.abs(d).max() < 1e-4
    assert abs(a["T"] - b["T"]) < 1e-3 * abs(b["T"])
