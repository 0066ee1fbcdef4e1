"""The reference's OWN driver against this repo's re-authored driver.

tests/ref_driver_harness.py runs /root/reference/src/start_simulation.py (with src/chemlab/*, converted to Python 3 in memory) on top of
chemlab_b200.espressopp -- `import espressopp` is the only thing that changes for it -- with the oracle as the backend.  The same
example then goes through `chemlab_b200.start_simulation` on the same backend, and every product of the two runs must be the same
file, byte for byte: energy CSV, end / before-reaction / whole configurations, output topology, bond / angle / dihedral tables,
reaction counters, topology-manager dumps, the parameter dump.  Only the wall-clock records differ (`_benchmark.csv/.pck`), and the
re-authored driver writes two extra files (`_state.dat`, `_bonds_chem_<k>.dat`).

What this pins: (1) the espressopp surface is a drop-in for the reference's driver code (row b of SURVEY 8): arg file, topology,
reactions incl. RestrictReaction, force-field assembly, thermostat, observers, the main loop and all writers run unmodified;
(2) the re-authored driver is the same program.  Python-2 dict order is switched off on our side for the comparison
(CHEMLAB_PY2_ORDER=0): the reference code runs under Python 3 here, where dicts keep insertion order.
Runs only where the reference tree is mounted (this container)."""
import filecmp
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree is not mounted")

CASES = {
    # 20 activated trimers, ATRP activator, neighbour-type changes, hooks.py, LJ by combination rule
    "atrp_lj": ["--run", "400", "--rng_seed", "42", "--start_ar", "200", "--energy_collect", "200"],
    # 36 tabulated pair potentials (15 conversion-mixed), tabulated bonds / angles / dihedrals: the 1000 LISTED dihedrals are static in
    # the reference although their types are dynamic (they are listed in the non-canonical orientation)
    "hyperbranched": ["--run", "500", "--rng_seed", "5"],
    # the optional observers and stop rules: tuple / type / type-state counters, monitor filter, Arrhenius rate update with its
    # rate file, .gro trajectory, two conversion criteria with eq_steps, stop_ar
    "atrp_lj+options": ["--run", "800", "--rng_seed", "42", "--start_ar", "200", "--energy_collect", "200", "--count_tuples", "True", "--count_types", "MA,FA",
                        "--count_types_state", "PL:1,FA:2", "--system_monitor_filter", "lj,count", "--rate_arrhenius", "True", "--gro_trj_collect", "400",
                        "--maximum_conversion", "MA:5000:6000,PL(1)+FA(7):2:2000", "--eq_steps", "16", "--stop_ar", "600"],
    # two reaction groups (their order!), 28 tabulated pair potentials, random partner selection, Akima reaction-bond tables
    "rim135": ["--run", "1000", "--rng_seed", "11", "--start_ar", "500", "--energy_collect", "500"],
    # network formation, intramolecular: 0 (molecule ids), state window up to 3
    "mf": ["--run", "1000", "--rng_seed", "3", "--start_ar", "0"],
    # p >= 1, nearest partner, virtual reactions
    "chain_growth_catalytic": ["--run", "1000", "--rng_seed", "3", "--start_ar", "500"],
    # RestrictReaction (connectivity map), CapForce, exclusion list from file, tabulated angles of 45,000 rows
    "dacron_restrict": ["--run", "200", "--rng_seed", "7", "--t_hybrid_bond", "0", "--int_step", "100", "--energy_collect", "100"],
}
OURS = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import chemlab_b200.espressopp._context as C\nfrom oracle.engine_adapter import OracleEngine\nC.Engine = OracleEngine\n"
        "from chemlab_b200 import start_simulation as S\nS.main(%r)\nprint('OURS_OK')\n")


def _products(d, example):
    out = {}
    for root, _, files in os.walk(d):
        for f in files:
            rel = os.path.relpath(os.path.join(root, f), d)
            if not os.path.exists(os.path.join(GOLD, example, rel)) and not rel.endswith((".pot", ".pyc")) and "__pycache__" not in rel:
                out[rel] = os.path.join(root, f)
    return out


@pytest.mark.parametrize("example", sorted(CASES))
def test_products_equal_those_of_the_reference_driver(example, tmp_path):
    sys.path.insert(0, ROOT)
    from chemlab_b200 import synthetic
    args = ["@params"] + CASES[example]
    example = example.split("+")[0]
    d_ref = synthetic.prepare_example(os.path.join(GOLD, example), str(tmp_path / "ref"), example)
    d_our = synthetic.prepare_example(os.path.join(GOLD, example), str(tmp_path / "ours"), example)
    os.makedirs(os.path.join(d_ref, "data"), exist_ok=True)        # the reference expects the directory of output_prefix to exist
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.Popen([sys.executable, os.path.join(HERE, "ref_driver_harness.py"), d_ref, d_ref] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    q = subprocess.Popen([sys.executable, "-c", OURS % (ROOT, HERE, args)], cwd=d_our, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         env=dict(env, CHEMLAB_PY2_ORDER="0"))
    po, pe = p.communicate(timeout=900)
    qo, qe = q.communicate(timeout=900)
    assert "REFERENCE_DRIVER_OK" in po and "Finished! Thanks!" in po, pe[-3000:]
    assert "OURS_OK" in qo, qe[-3000:]
    a, b = _products(d_ref, example), _products(d_our, example)
    timing = {k for k in a if k.endswith(("_benchmark.csv", "_benchmark.pck"))}
    assert set(a) <= set(b) and len(a) >= 15
    assert {os.path.basename(k).split("_", 2)[-1] for k in set(b) - set(a)} <= {"state.dat", "bonds_chem_0.dat", "bonds_chem_1.dat"} or \
        all(k.endswith(("_state.dat", "_bonds_chem_0.dat", "_bonds_chem_1.dat")) for k in set(b) - set(a))
    different = sorted(k for k in set(a) - timing if not filecmp.cmp(a[k], b[k], shallow=False))
    assert not different, different
    # the run did something: the energy file has rows, and the reaction lists are not empty in the cases that react within the window
    energy = [k for k in a if "_energy_" in k]
    assert len(energy) == 1 and len(open(a[energy[0]]).read().splitlines()) >= 3


def test_topology_preparation_equals_the_reference_code_on_every_shipped_topology(tmp_path):
    """GromacsTopology(...).read() of the reference (src/chemlab/gromacs_topology.py, converted in memory) against ours on every
    shipped .top: type ids, the replicated atoms with their parameters, bonds / angles / dihedrals / pairs with the parameters of
    their lines, the parameter dictionaries keyed by type-id tuples, and the generated exclusions.  In a child process (the
    reference module needs `espressopp` to be this repo's surface)."""
    code = r'''
import os, sys, glob, tempfile, io, contextlib
os.environ["CHEMLAB_PY2_ORDER"] = "0"
sys.path.insert(0, %r); sys.path.insert(0, %r)
import ref_driver_harness as H
tmp = H.build(tempfile.mkdtemp(prefix="chemlab_ref_topology_"))
sys.path.insert(0, os.path.join(tmp, "chemlab")); sys.path.insert(0, tmp)
import chemlab_b200.espressopp as es
sys.modules["espressopp"] = es
for sub in ("analysis", "integrator", "interaction", "storage", "bc", "esutil", "io", "tools"):
    sys.modules["espressopp." + sub] = getattr(es, sub)
import gromacs_topology as ref_gt
from chemlab_b200.chemlab import gromacs_topology as our_gt
def norm(o):
    if isinstance(o, dict): return {str(k): norm(v) for k, v in o.items()}
    if isinstance(o, (set, frozenset)): return sorted(norm(v) for v in o)
    if isinstance(o, (list, tuple)): return [norm(v) for v in o]
    if isinstance(o, float): return round(o, 12)
    return o
n = 0
for p in sorted(glob.glob("/root/reference/examples/**/*.top", recursive=True)) + ["/root/reference/src/tests/topol.top"]:
    if p.endswith("atrp_activator/topol.top"):
        continue                      # includes idd.itp, which the reference does not ship
    os.chdir(os.path.dirname(p))
    with contextlib.redirect_stdout(io.StringIO()):
        a = ref_gt.GromacsTopology(os.path.basename(p)); a.read()
        b = our_gt.GromacsTopology(os.path.basename(p), generate_exclusions=True).read()
    for attr in ("atomsym_atomtype", "used_atomsym_atomtype", "atoms", "bonds", "angles", "dihedrals", "pairs", "bondparams", "angleparams", "dihedralparams", "exclusions"):
        x, y = getattr(a, attr), getattr(b, attr)
        if attr == "exclusions":
            x, y = {tuple(sorted(t)) for t in x}, {tuple(sorted(t)) for t in y}
        assert norm(x) == norm(y), (p, attr)
    n += 1
assert n >= 15
import shutil; shutil.rmtree(tmp, ignore_errors=True)
print("TOPOLOGY_OK", n)
''' % (ROOT, HERE)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "TOPOLOGY_OK" in p.stdout, p.stderr[-3000:]
