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
