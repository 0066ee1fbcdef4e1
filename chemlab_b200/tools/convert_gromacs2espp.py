"""python -m chemlab_b200.tools.convert_gromacs2espp in.xvg out.pot -- GROMACS table -> ESPResSo++ `r e f` table
(tools/convert_gromacs2espp.py:112-123).  The conversion itself is espressopp.tools.convert.gromacs.convertTable, which
reproduces the shipped .xvg/.pot pairs byte for byte (tests/test_golden_cpu.py)."""
import argparse

from ..espressopp.tools.convert.gromacs import convertTable


def _args():
    p = argparse.ArgumentParser(description="Convert a GROMACS .xvg table to the .pot format")
    p.add_argument("in_file")
    p.add_argument("out_file")
    return p


def main(argv=None):
    a = _args().parse_args(argv)
    convertTable(a.in_file, a.out_file)


if __name__ == "__main__":
    main()
