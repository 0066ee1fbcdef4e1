"""Known-answer tests against the artefacts the reference DOES ship: table conversion .xvg -> .pot
(byte-exact), the .pot reader feeding clb_add_table, Philox.  Fixtures under tests/golden/ were copied out of
/root/reference by tests/golden/make_golden.py (committed) because the reference tree does not travel to the GPU box."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _pairs():
    out = []
    for f in sorted(os.listdir(GOLD)):
        if f.endswith(".xvg") and os.path.exists(os.path.join(GOLD, f[:-4] + ".pot")):
            out.append(f[:-4])
    return out


@pytest.mark.parametrize("stem", _pairs())
def test_convert_table_reproduces_shipped_pot_byte_for_byte(stem, tmp_path):
    from chemlab_b200.espressopp.tools.convert import gromacs
    out = tmp_path / (stem + ".pot")
    gromacs.convertTable(os.path.join(GOLD, stem + ".xvg"), str(out))
    assert out.read_bytes() == open(os.path.join(GOLD, stem + ".pot"), "rb").read()


@pytest.mark.parametrize("stem", _pairs())
def test_pot_reader_gives_uniform_grid(stem):
    from chemlab_b200.espressopp.interaction import read_pot
    r, e, f = read_pot(os.path.join(GOLD, stem + ".pot"))
    assert len(r) == len(e) == len(f) > 10
    d = np.diff(r)
    assert np.allclose(d, d[0], rtol=1e-5, atol=1e-9)


def test_oracle_table_matches_numpy_interp():
    from chemlab_b200.espressopp.interaction import read_pot
    from oracle import pyoracle
    stem = _pairs()[0]
    r, e, f = read_pot(os.path.join(GOLD, stem + ".pot"))
    o = pyoracle.Oracle(1, [10, 10, 10], 1.0, 0.1)
    t = o.add_table(r, e, f, 1)
    xs = np.linspace(r[0], r[-1] * 0.999, 57)
    for x in xs:
        ee, ff, bad = o.table_eval(t, float(x))
        assert not bad
        assert abs(ee - np.interp(x, r, e)) <= 1e-9 * max(1, abs(ee))
        assert abs(ff - np.interp(x, r, f)) <= 1e-9 * max(1, abs(ff))


def test_table_tools_command_line(tmp_path, monkeypatch):
    """The three table tools of the reference's tools/ directory as `python -m chemlab_b200.tools.<name>`: the converter's command
    line reproduces a shipped .pot byte for byte; fix_table repairs zero forces at the table ends in place; mix_table writes
    x*tab1 + (1-x)*tab2 (and the geometric variant) for every func-9 row of the topology (tools/mix_table.py:107-123)."""
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    env = dict(os.environ, PYTHONPATH=root)
    out = tmp_path / "b0.pot"
    subprocess.check_call([sys.executable, "-m", "chemlab_b200.tools.convert_gromacs2espp", os.path.join(GOLD, "table_b0.xvg"), str(out)], env=env)
    assert out.read_bytes() == open(os.path.join(GOLD, "table_b0.pot"), "rb").read()
    # fix_table
    from chemlab_b200.tools import fix_table, mix_table
    t = np.column_stack([np.arange(1, 6) * 0.1, np.arange(5) * 1.0, [0.0, 2.0, 3.0, 4.0, 0.0]])
    p = tmp_path / "t.pot"
    np.savetxt(p, t)
    fix_table.main([str(p)])
    d = np.loadtxt(p)
    assert d[0, 2] == 2.0 and d[-1, 2] == 4.0 and (d[1:-1] == t[1:-1]).all()
    # mix_table on two shipped-format tables of different length
    monkeypatch.chdir(tmp_path)
    xa = np.loadtxt(os.path.join(GOLD, "table_A_A.xvg"))
    xb = xa[:-7].copy(); xb[:, 3:] *= 0.5
    np.savetxt("table_MO_MO.xvg", xa); np.savetxt("table_PO_PO.xvg", xb)
    open("topol.top", "w").write("[ defaults ]\n1 1 no 1.0 1.0\n\n[ atomtypes ]\nMO 1.0 0.0 A 1.0 1.0\nPO 1.0 0.0 A 1.0 1.0\n\n"
                                 "[ nonbond_params ]\nMO PO 9 tabA tabB PO 100 0.0 0.5\n\n[ moleculetype ]\nM 1\n\n[ atoms ]\n1 MO 1 M A1 1 0.0 1.0\n\n"
                                 "[ system ]\nx\n\n[ molecules ]\nM 1\n")
    assert mix_table.main(["--scaling", "0.25"]) == ["table_tabB_tabA.pot"]
    m = np.loadtxt("table_tabB_tabA.pot")
    ra, rb = mix_table.xvg_to_ref(xa), mix_table.xvg_to_ref(xb)
    assert len(m) == len(xb) and np.allclose(m[:, 1], 0.25 * ra[:len(xb), 1] + 0.75 * rb[:, 1], rtol=1e-9, atol=1e-12)
    assert np.allclose(m[:, 2], (0.25 + 0.75 * 0.5) * ra[:len(xb), 2], rtol=1e-9, atol=1e-12)
    g = mix_table.mix_geometric(np.array([[0.1, 4.0, 1.0]]), np.array([[0.1, 9.0, 2.0]]), 0.5, 0.0)
    assert np.allclose(g[0], [0.1, 2.0 + 3.0, 0.5 * 0.5 * 1.0 + 0.5 * (1.0 / 3.0) * 2.0])


@pytest.mark.parametrize("rel,exact", [("dacron/conf.gro", True), ("dacron_restrict/conf.gro", True), ("hyperbranched/conf.gro", True),
                                       ("chain_growth_catalytic/conf.gro", True), ("rim135/cg_conf.gro", False), ("pccg_lj/conf.gro", False)])
def test_gro_writer_reproduces_the_files_the_reference_wrote(rel, exact, tmp_path):
    """These shipped coordinate files were written by the reference's own GROFile.write (title 'XXX of molecules', '%d' atom count,
    box line '%f %f %f': src/chemlab/files_io.py:216-257): reading them and writing them again must give the same bytes
    (two of them were hand-edited at the very end of the file: equal up to trailing newlines)."""
    from chemlab_b200.chemlab.files_io import GROFile
    src = os.path.join(GOLD, rel)
    g = GROFile(src); g.read()
    out = GROFile(str(tmp_path / "out.gro"))
    out.box, out.title, out.atoms = g.box, g.title, g.atoms
    out.write()
    a, b = open(src, "rb").read(), open(str(tmp_path / "out.gro"), "rb").read()
    assert (a == b) if exact else (a.rstrip(b"\n") == b.rstrip(b"\n"))
    assert b.endswith(b"\n")
