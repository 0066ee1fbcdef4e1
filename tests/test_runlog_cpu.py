"""Known answers from the one run log the reference ships: examples/atrp_lj/single -- the stdout of
`start_simulation.py` on the 6000-bead tabulated ATRP system of examples/atrp_activator (box 13.40248, rng seed 3036).  The same
set-up is driven through this repo's driver (oracle backend, no GPU) and every fact the log states about the set-up is compared:
cell grid, density, exclusion count, type ids, dynamic types, the reaction type changes, the neighbour-change rules, the
ATRPActivator centres, the registered angle type tuples, the integrator step, the collection interval, the monitored labels.

The log predates three later additions to the example directory, which are removed from the scratch copy: the initiator molecule
(`#include "idd.itp"`, itself missing from the reference; its [ atomtypes ] row `I`; [ molecules ] EGD 1998 + IDD 3 -> EGD 2000 as the
log says), the
truncated `I I` row of [ nonbond_params ] (the reference's own parser raises on it) and the `I:I` dissociation reaction.
table_a0 is a missing blob (stand-in: table_a1); the Python-2 hook file is replaced by the Python-3 one of tests/golden/atrp_lj.
Runs only where the reference tree is mounted (this container)."""
import contextlib
import io
import os
import re
import shutil

import numpy as np
import pytest

REF = "/root/reference/examples"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference examples are not mounted")


@pytest.fixture(scope="module")
def logged_run(tmp_path_factory):
    import sys
    sys.path.insert(0, HERE)
    import chemlab_b200.espressopp._context as C
    from oracle.engine_adapter import OracleEngine
    from chemlab_b200 import start_simulation as S
    d = str(tmp_path_factory.mktemp("runlog") / "atrp_activator")
    shutil.copytree(os.path.join(REF, "atrp_activator"), d)
    for root, _, files in os.walk(d):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    os.chmod(d, 0o755)

    def edit(name, fn):
        p = os.path.join(d, name)
        s = open(p).read()
        t = fn(s)
        assert t != s, name
        open(p, "w").write(t)
    edit("topol.top", lambda s: re.sub(r"\n  I  .*", "", s.replace('#include "idd.itp"\n', "").replace("EGD             1998", "EGD             2000").replace("IDD             3\n", "")))
    edit("ffnb.itp", lambda s: re.sub(r"\n I    I\s*$", "\n", s))
    edit("atrp.cfg", lambda s: s[:s.index("[reaction_rev]")])
    for ext in ("pot", "xvg"):
        shutil.copy(os.path.join(d, "table_a1." + ext), os.path.join(d, "table_a0." + ext))
    shutil.copy(os.path.join(HERE, "golden", "atrp_lj", "hooks.py"), os.path.join(d, "hooks.py"))
    os.remove(os.path.join(d, "exclusion_topol.list"))          # written for EGD 1998 + IDD 3; the logged run generated its own 6000
    os.makedirs(os.path.join(d, "data"), exist_ok=True)
    cwd = os.getcwd()
    real = C.Engine
    C.Engine = OracleEngine
    os.chdir(d)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            r = S.main(["@params", "--rng_seed", "3036", "--run", "200", "--gen_velocity", "True"])
    finally:
        C.Engine = real
        os.chdir(cwd)
    return r, buf.getvalue(), open(os.path.join(REF, "atrp_lj", "single")).read()


def _find(rx, text, cast=str):
    m = re.search(rx, text, re.M)
    assert m, rx
    return cast(m.group(1))


def test_geometry_density_and_counts(logged_run):
    r, out, log = logged_run
    assert _find(r"^Cell grid: (\(.*?\))", log) == _find(r"^Cell grid: (\(.*?\))", out) == "(6, 6, 6)"
    assert abs(_find(r"^Density: ([0-9.]+) kg", out, float) - _find(r"^Density: ([0-9.]+) kg", log, float)) < 1e-8      # 273.445308845
    assert _find(r"^Excluded pairs from LJ interaction: (\d+)", out, int) == _find(r"^Excluded pairs from LJ interaction: (\d+)", log, int) == 6000
    assert _find(r"^(Reads \d+ particles with properties .*)$", out) == _find(r"^(Reads \d+ particles with properties .*)$", log)
    assert _find(r"^(Using kB=.*)$", log) == "Using kB=0.0083144621 and mass-factor=1.6605402"
    assert _find(r"^Boltzmann constant: (.*)$", out) == _find(r"^Boltzmann constant: (.*)$", log) == "0.0083144621"
    assert _find(r"^Skin: (.*)$", out, float) == _find(r"^Skin: (.*)$", log, float) == 0.1
    assert _find(r"distribution T=343.0 \(([0-9.]+)\)", out, float) == _find(r"distribution T=343.0 \(([0-9.]+)\)", log, float)
    gt = r["topology"]
    assert (len(gt.bonds), len(gt.angles), len(gt.dihedrals)) == tuple(_find(r"^%s: (\d+)" % k, log, int) for k in ("Bonds", "Angles", "Dihedrals"))


def test_type_ids_and_dynamic_types(logged_run):
    r, out, log = logged_run
    table = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"^([A-Z]{2})\s+(\d+)\s*$", log, re.M))
    assert table == {"MA": 0, "ML": 1, "DA": 2, "FA": 3, "PA": 4, "RA": 5, "PL": 6}
    ids = r["topology"].atomsym_atomtype
    assert {k: ids[k] for k in table} == table
    dyn = set(int(x) for x in _find(r"^Dynamic type ids: set\(\[(.*?)\]\)", log).split(","))
    ours = set(int(x) for x in re.findall(r"\d+", _find(r"^Dynamic type ids: (.*)$", out)))
    assert ours == dyn


def test_reaction_setup_matches_the_log(logged_run):
    r, out, log = logged_run
    import chemlab_b200.espressopp as es
    ids = r["topology"].atomsym_atomtype
    # "Setup reaction: FA(3)-MA(0)" followed by its "Reaction: FA-MA, change type a->b" lines
    logged, cur = [], None
    for line in log.splitlines():
        m = re.match(r"Setup reaction: (\w+)\((\d+)\)-(\w+)\((\d+)\)", line)
        if m:
            cur = {"t": (int(m.group(2)), int(m.group(4))), "chg": set()}
            logged.append(cur)
        m = re.match(r"Reaction: \w+-\w+, change type (\d+)->(\d+)", line)
        if m:
            cur["chg"].add((int(m.group(1)), int(m.group(2))))
    assert len(logged) == 4 == len(r["reactions"])
    for want, reaction in zip(logged, r["reactions"]):
        assert (reaction.type_1, reaction.type_2) == want["t"]
        got = {(old, int(p.type)) for pp, _ in reaction._post if type(pp) is es.integrator.PostProcessChangeProperty for old, p, lvl in pp._rules}
        assert got == want["chg"], (want, got)
        nb = {(old, int(p.type), lvl) for pp, _ in reaction._post if isinstance(pp, es.integrator.PostProcessChangeNeighboursProperty)
              for old, p, lvl in pp._rules}
        # "Change property MA->PA nb=2", "Change property ML->PL nb=1"
        want_nb = {(ids[m.group(1)], ids[m.group(2)], int(m.group(3))) for m in re.finditer(r"^Change property (\w+)->(\w+) nb=(\d+)", log, re.M)}
        # the shipped atrp.cfg has since gained a third rule, PL:1->PL(state=1)
        assert len(want_nb) == 2 and want_nb <= nb and nb - want_nb == {(ids["PL"], ids["PL"], 1)}
    assert _find(r"^Change integrator step to (\d+)", log, int) == _find(r"x integrator\.run\((\d+)\)", out, int) == 200
    assert _find(r"collect data every (\d+) steps", log, int) == _find(r"collect data every (\d+) steps", out, int) == 200


def test_atrp_activator_centres_match_the_log(logged_run):
    r, out, log = logged_run
    import chemlab_b200.espressopp as es
    ids = r["topology"].atomsym_atomtype
    act = [x for x in r["integrator"]._extensions if isinstance(x, es.integrator.ATRPActivator)]
    assert len(act) == 1
    act = act[0]
    m = re.search(r"ATRPActivator\.interval=(\d+) num_part=(\d+)", log)
    assert (act.interval, act.num_particles) == (int(m.group(1)), int(m.group(2))) == (200, 1000)
    want = []
    for m in re.finditer(r"^ATRPActivator: added (\w+)\((\d+),(\w+)\)->(\w+)\((-?\d+)\) state=(\d+) is_activator=(\w+) delta_state=(-?\d+)", log, re.M):
        assert m.group(2) == m.group(6) and m.group(5) == m.group(8)
        want.append((ids[m.group(1)], int(m.group(6)), m.group(7) == "True", ids[m.group(4)], int(m.group(8))))
    got = [(t, s, a, int(p.type), ds) for t, s, a, p, ds in act._centers]
    assert len(want) == 4 and got == want


def test_registered_angles_and_monitored_labels_match_the_log(logged_run):
    r, out, log = logged_run
    want = {tuple(int(x) for x in m.group(1).split(",")) for m in re.finditer(r"^Register angles for type: \((.*?)\)", log, re.M)}
    got = {t for _, t in r["system"].topology_manager._triplets}
    assert len(want) == 19 and got == want
    logged_labels = re.findall(r"^System analysis: adding (\S+)", log, re.M)
    assert logged_labels == sorted(logged_labels)                        # the reference adds them sorted by label (:471)
    ours = [l for l in re.findall(r"(\S+)=", _find(r"^(step 0: .*)$", out)) if l not in ("T", "Ekin") and not l.startswith(("cr_", "count_"))]
    assert ours == sorted(ours)
    # the reaction list was called fpl_<group> when the log was written, chem_fpl_<group> in the shipped source (reaction_setup.py:467)
    assert sorted(l.replace("chem_fpl_", "fpl_") for l in ours) == logged_labels


def test_force_field_assembly_matches_the_log(logged_run):
    """Which table every type pair reads, which pairs are conversion-mixed, and the table number of every angle type tuple."""
    r, out, log = logged_run
    assert _find(r"^Number of non-bonded type pairs: (\d+)", out, int) == _find(r"^Number of non-bonded type pairs: (\d+)", log, int) == 28
    tab = lambda text: {(frozenset(m.group(1).split("-")), m.group(2)) for m in re.finditer(r"^Set tab potential (\S+): (\S+)", text, re.M)}
    assert tab(out) == tab(log) and len(tab(log)) == 24
    mixed = lambda text: {frozenset(int(x) for x in m.group(1).split("-")) for m in re.finditer(r"^Set mixed tabulated potential (\d+-\d+) ", text, re.M)}
    assert mixed(out) == mixed(log) == {frozenset((1, 6)), frozenset((0, 4))}
    gt = r["topology"]
    n = 0
    for m in re.finditer(r"^\((\d+), (\d+), (\d+)\) \{'params': \['(\d+)', '([0-9.]+)'\], 'func': (\d+)\}", log, re.M):
        key = tuple(int(m.group(k)) for k in (1, 2, 3))
        p = gt.angleparams.get(key) or gt.angleparams.get(key[::-1])
        assert p is not None and int(p["func"]) == int(m.group(6)) and [str(x) for x in p["params"]] == [m.group(4), m.group(5)], (key, p)
        n += 1
    assert n == 19
