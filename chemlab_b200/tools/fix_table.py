"""python -m chemlab_b200.tools.fix_table table.pot -- repair the two end nodes of a tabulated potential in place.

Tables produced by numerical differentiation carry a force of exactly zero at the first and/or the last node (the derivative is not
defined there).  Such a node takes the force of its inner neighbour; every other row is left as it is.  Same job as the reference's
tools/fix_table.py:20-31."""
import sys

import numpy as np

FORCE = 2      # columns: r, energy, force


def fix_table(path):
    table = np.loadtxt(path)
    for end, inner in ((0, 1), (-1, -2)):
        if table[end, FORCE] == 0.0:
            table[end, FORCE] = table[inner, FORCE]
    np.savetxt(path, table)
    return table


def main(argv=None):
    argv = sys.argv[1:] if argv is None else list(argv)
    if len(argv) != 1:
        raise SystemExit("usage: python -m chemlab_b200.tools.fix_table <table file>")
    fix_table(argv[0])


if __name__ == "__main__":
    main()
