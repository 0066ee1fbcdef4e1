"""python -m chemlab_b200.tools.fix_table table.pot -- a force column that starts or ends with an exact zero takes the value of
its neighbour, in place (tools/fix_table.py:20-31: tabulated bonded potentials whose finite-difference force is undefined at
the first/last node)."""
import sys

import numpy as np


def fix_table(path):
    d = np.loadtxt(path)
    if d[0][2] == 0.0:
        d[0][2] = d[1][2]
    if d[-1][2] == 0.0:
        d[-1][2] = d[-2][2]
    np.savetxt(path, d)
    return d


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        raise SystemExit("usage: python -m chemlab_b200.tools.fix_table <table>")
    fix_table(argv[0])


if __name__ == "__main__":
    main()
