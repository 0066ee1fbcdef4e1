"""tools.convert.gromacs.convertTable: GROMACS .xvg -> ESPResSo++ .pot (behaviour of tools/convert_gromacs2espp.py:28-110,
pinned byte-for-byte by the shipped table pairs, see tests/test_golden_cpu.py)."""
import math
import re

_KINDS = (("bond", re.compile(r".*_b[0-9]+.*")), ("angle", re.compile(r".*_a[0-9]+.*")), ("dihedral", re.compile(r".*_d[0-9]+.*")))
_ROW = "%15.8g %15.8g %15.8g\n"


def table_kind(filename):
    for kind, rx in _KINDS:
        if rx.match(filename):
            return kind
    return "nonbonded"


def convertTable(gro_in_file, esp_out_file, sigma=1.0, epsilon=1.0, c6=1.0, c12=1.0):
    kind = table_kind(gro_in_file)
    out = []
    with open(gro_in_file) as fin:
        for line in fin:
            if line.startswith("#"):
                continue
            c = line.split()
            if kind == "nonbonded":
                # columns: r f f' g g' h h' ; electrostatics (f, f') ignored
                r = float(c[0]) / sigma
                e = (c6 * float(c[3]) + c12 * float(c[5])) / epsilon
                f = (c6 * float(c[4]) + c12 * float(c[6])) * sigma / epsilon
                keep = r != 0
            else:
                x, e_raw, f_raw = float(c[0]), float(c[1]), float(c[2])
                if kind == "bond":
                    r = x / sigma
                    keep = r != 0
                else:
                    r = math.radians(x)
                    f_raw = f_raw * 180 / math.pi
                    keep = (0 < r <= math.pi) if kind == "angle" else (-math.pi <= r <= math.pi)
                e = e_raw / epsilon
                f = f_raw * sigma / epsilon
            if keep:
                out.append(_ROW % (r, e, f))
    with open(esp_out_file, "w") as fout:
        fout.writelines(out)
