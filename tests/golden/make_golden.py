"""Copies the small known-answer artefacts out of /root/reference into tests/golden/ (run once, here; the reference
tree does not exist on the GPU box).  Only DATA files are copied (tables, topology fixtures, exclusion lists)."""
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = [
    # (.xvg, .pot) pairs reproduced byte-for-byte by tools/convert_gromacs2espp.py (SURVEY 8c [PROBE])
    "examples/hyperbranched/table_b0.xvg", "examples/hyperbranched/table_b0.pot",
    "examples/dacron/rev_with_water/test_3/table_b2.xvg", "examples/dacron/rev_with_water/test_3/table_b2.pot",
    "examples/dacron/rev_with_water/test_3/table_A_A.xvg", "examples/dacron/rev_with_water/test_3/table_A_A.pot",
    # parser fixtures of the reference's own unit tests
    "src/tests/topol.top", "src/tests/diol_cg.itp", "src/tests/ter_cg.itp",
    # atrp_lj inputs (config 1) + the exclusion list the reference wrote for it
    "examples/atrp_lj/topol.top", "examples/atrp_lj/ffnb.itp", "examples/atrp_lj/atrp.cfg", "examples/atrp_lj/params",
    "examples/atrp_lj/exclusion_topol.list", "examples/atrp_lj/conf.gro",
    # chain_growth_catalytic inputs: the shipped RNG-free reaction config (p = rate*dt*interval = 2.5 >= 1, nearest partner)
    "examples/chain_growth_catalytic/topol.top", "examples/chain_growth_catalytic/reaction.cfg",
    "examples/chain_growth_catalytic/params", "examples/chain_growth_catalytic/conf.gro",
]
for f in FILES:
    src = os.path.join(REF, f)
    sub = "atrp_lj" if "atrp_lj" in f else ("chain_growth_catalytic" if "chain_growth_catalytic" in f else ("parser" if f.startswith("src/tests") else ""))
    dst_dir = os.path.join(HERE, sub)
    os.makedirs(dst_dir, exist_ok=True)
    shutil.copy(src, os.path.join(dst_dir, os.path.basename(f)))
    print("copied", f, os.path.getsize(src))
