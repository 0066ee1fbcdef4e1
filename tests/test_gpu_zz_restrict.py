"""Added after the last full GPU run of round 2 (the file sorts last on purpose).

RestrictReaction on the device (reaction_setup.py:74-75,115-126; examples/dacron/restrict): a reaction that
carries a connectivity map takes candidates only among the pairs named in it.  Candidate rows, events, bond lists and the
resulting types/states must equal the oracle's bit-exactly; the map can be replaced between passes; an empty map switches
the reaction off; other reactions of the same pass are not affected.  (The file sorts last on purpose: it was added after the
round's last full GPU run.)"""
import numpy as np
import pytest

import clb_testutil as util  # noqa: F401
from test_gpu_reactions import _add_both, _compare_state, _reactive_pair

pytestmark = pytest.mark.gpu


def _key(rows):
    return {(int(min(a, b)), int(max(a, b))) for a, b in np.asarray(rows)[:, :2]}


def test_restrict_reaction_candidates_are_bit_exact():
    m, P, h = _reactive_pair(nearest=0)
    r0 = _add_both(P, 0, 0, 1, 1, 1, 2, 1, 2, 1e6, 1.2, h["rl"], intramolecular=1, intraresidual=0)      # A(1,2)+A(1,2), to be restricted
    r1 = _add_both(P, 0, 1, 1, 1, 1, 2, 0, 1, 1e6, 1.7, h["rl"], intramolecular=1, intraresidual=1)      # A(1,2)+L(0,1), unrestricted (A-L of different trimers sit on lattice diagonals, 1.5)
    pairs = P.o.pairs()       # any pair set will do as a map; the two Verlet lists need not be equal here (different rebuild times)
    # the map: every second Verlet pair, written in reverse order and partly twice (define_connection is order-free)
    cmap = pairs[::2]
    P.both("reaction_define_connections", r0, np.concatenate([cmap[:, ::-1], cmap[:100]]))
    na, nb = P.e.react_now(), P.o.react()
    ca, da = P.e.last_candidates(); cb, db = P.o.candidates()
    assert len(ca) == len(cb) and (ca == cb).all()
    assert np.allclose(da, db, rtol=1e-12)
    allowed = _key(cmap)
    c0 = ca[ca[:, 2] == r0]
    assert len(c0) > 20 and _key(c0) <= allowed                       # restricted reaction: only pairs of the map
    assert (ca[:, 2] == r1).sum() > 0 and not _key(ca[ca[:, 2] == r1]) <= allowed   # the other reaction is not restricted
    assert na == nb > 10
    _compare_state(P, h)
    # replace the map by the other half of the pairs: the next pass sees those and none of the first half
    cmap2 = pairs[1::2]
    P.both("reaction_define_connections", r0, cmap2)
    assert P.e.react_now() == P.o.react()
    ca, _ = P.e.last_candidates(); cb, _ = P.o.candidates()
    assert len(ca) == len(cb) and (ca == cb).all()
    assert _key(ca[ca[:, 2] == r0]) <= _key(cmap2)
    _compare_state(P, h)
    # an empty map leaves a restricted reaction without candidates
    P.both("reaction_define_connections", r0, np.zeros((0, 2), np.int64))
    assert P.e.react_now() == P.o.react()
    ca, _ = P.e.last_candidates(); cb, _ = P.o.candidates()
    assert len(ca) == len(cb) and (ca == cb).all() and (ca[:, 2] == r0).sum() == 0
    _compare_state(P, h)
    with pytest.raises(Exception):
        P.e.reaction_define_connections(r0, np.array([[0, 10 ** 9]]))          # unknown particle id
    P.close()


def test_dacron_restrict_driver_gpu_matches_oracle(tmp_path):
    """examples/dacron/restrict through the chemlab driver on the GPU engine and on the oracle: the same bonds (all of them lines
    of connections.list), types, states and masses."""
    from test_driver_cpu import run_dacron_restrict
    a = run_dacron_restrict(str(tmp_path), "gpu", 600)
    b = run_dacron_restrict(str(tmp_path), "oracle", 600)
    assert a["steps"] == b["steps"] == 600
    assert len(b["bonds"]) >= 1 and a["bonds"].shape == b["bonds"].shape
    srt = lambda x: x[np.lexsort((x[:, 1], x[:, 0]))]
    assert (srt(np.sort(a["bonds"], 1)) == srt(np.sort(b["bonds"], 1))).all()
    assert all(tuple(sorted(x)) in a["conn"] for x in a["bonds"].tolist())
    assert (a["g"]["type"] == b["g"]["type"]).all() and (a["g"]["state"] == b["g"]["state"]).all()
    assert np.array_equal(a["g"]["mass"], b["g"]["mass"])


def test_mf_driver_gpu_matches_oracle(tmp_path):
    """examples/mf/espp_cg_1 as shipped (one bead type, A(0,3) + A(0,3) -> A(1):A(1), intramolecular: 0 -> molecule ids merge
    with every bond and later passes must see them): GPU run == oracle run, bond for bond."""
    from test_driver_cpu import run_mf
    # five passes within 2 ps: the engine's 2^-32 L position lattice against the oracle's doubles has not grown beyond 1e-6 nm yet
    a = run_mf(str(tmp_path), "gpu", 1000, interval=200)
    b = run_mf(str(tmp_path), "oracle", 1000, interval=200)
    srt = lambda x: x[np.lexsort((x[:, 1], x[:, 0]))]
    assert len(b["bonds"]) > 5 and a["bonds"].shape == b["bonds"].shape
    assert (srt(np.sort(a["bonds"], 1)) == srt(np.sort(b["bonds"], 1))).all()
    assert (a["g"]["state"] == b["g"]["state"]).all() and (a["g"]["type"] == b["g"]["type"]).all()


def test_pccg_lj_driver_gpu_matches_oracle(tmp_path):
    """examples/pccg_lj/chemical_reactions (15,200 beads: pair-specific LJ, FENE + LJ monomer and reaction bonds, Cosine angles from the
    TopologyManager, ATRPActivator, the five user hooks): GPU run == oracle run."""
    from test_driver_cpu import run_pccg_lj
    a = run_pccg_lj(str(tmp_path), "gpu", 600)
    b = run_pccg_lj(str(tmp_path), "oracle", 600)
    srt = lambda x: x[np.lexsort((x[:, 1], x[:, 0]))]
    assert len(b["bonds"]) >= 3 and a["bonds"].shape == b["bonds"].shape
    assert (srt(np.sort(a["bonds"], 1)) == srt(np.sort(b["bonds"], 1))).all()
    assert (a["g"]["state"] == b["g"]["state"]).all() and (a["g"]["type"] == b["g"]["type"]).all()
    assert a["names"] == b["names"]
