"""python -m chemlab_b200.tools.mix_table [--top topol.top] [--scaling x] [--mix_type 0|1] [--constant c] -- static mixtures of two
non-bonded tables for every `[ nonbond_params ]` row with func 9 (tools/mix_table.py:39-123): for the row `t1 t2 9 name_a name_b ...`
table_<t1>_<t1>.xvg (monomer) and table_<t2>_<t2>.xvg (polymer) are converted (e = g + h, f = g' + h'; r = 0 kept, unlike
convertTable) and mixed into table_<name_b>_<name_a>.pot,

    arithmetic (0):  y = x y1 + (1 - x) y2                       for energy and force
    geometric  (1):  e = (e1 + c)^x + (e2 + c)^(1-x) - c,  f = x (e1 + c)^(x-1) f1 + (1 - x) (e2 + c)^(-x) f2

(the geometric expressions are the reference's as written).  Tables of different length are cut to the shorter one."""
import argparse
import datetime

import numpy as np


def xvg_to_ref(xvg):
    """columns r f f' g g' h h' -> r, g + h, g' + h'"""
    xvg = np.asarray(xvg, float)
    return np.column_stack([xvg[:, 0], xvg[:, 3] + xvg[:, 5], xvg[:, 4] + xvg[:, 6]])


def _common(tab1, tab2):
    n = min(len(tab1), len(tab2))
    if n == 0:
        raise RuntimeError("The length of output table is zero")
    if len(tab1) != len(tab2) and (tab1[:n, 0] != tab2[:n, 0]).all():
        raise RuntimeError("Both r columns should be the same")
    return n


def mix_arithmetic(tab1, tab2, coupling):
    n = _common(tab1, tab2)
    out = np.array(tab1[:n], float)
    out[:, 1:3] = coupling * tab1[:n, 1:3] + (1.0 - coupling) * tab2[:n, 1:3]
    return out


def mix_geometric(tab1, tab2, coupling, constant):
    n = _common(tab1, tab2)
    out = np.array(tab1[:n], float)
    e1, e2, f1, f2 = tab1[:n, 1], tab2[:n, 1], tab1[:n, 2], tab2[:n, 2]
    out[:, 1] = np.power(e1 + constant, coupling) + np.power(e2 + constant, 1.0 - coupling) - constant
    out[:, 2] = coupling * np.power(e1 + constant, coupling - 1.0) * f1 + (1.0 - coupling) * np.power(e2 + constant, -coupling) * f2
    return out


def main(argv=None):
    p = argparse.ArgumentParser("Mix table")
    p.add_argument("--top", default="topol.top")
    p.add_argument("--scaling", type=float, default=0.5, help="scaling factor x")
    p.add_argument("--constant", type=float, default=0.0, help="constant, for the geometric type")
    p.add_argument("--mix_type", type=int, default=0, choices=[0, 1], help="0 arithmetic, 1 geometric")
    a = p.parse_args(argv)
    from ..chemlab.gromacs_topology import GromacsTopology
    topol = GromacsTopology(a.top).read()
    written = []
    for (t1, t2), params in topol.gt.nonbond_params.items():
        if params["func"] != 9:
            continue
        mono = xvg_to_ref(np.loadtxt("table_%s_%s.xvg" % (t1, t1)))
        poly = xvg_to_ref(np.loadtxt("table_%s_%s.xvg" % (t2, t2)))
        out_name = "table_%s_%s.pot" % (params["params"][1], params["params"][0])
        mixed = mix_arithmetic(mono, poly, a.scaling) if a.mix_type == 0 else mix_geometric(mono, poly, a.scaling, a.constant)
        np.savetxt(out_name, mixed, header="Mixed of %s and %s at %s" % (t1, t2, datetime.datetime.now()), fmt="%2.9e")
        print("Saved %s" % out_name)
        written.append(out_name)
    return written


if __name__ == "__main__":
    main()
