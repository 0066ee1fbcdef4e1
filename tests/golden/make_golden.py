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


# rim135 (config 5 base system): topology, coordinates, arg-file, reaction config as shipped; the 28 non-bonded, 5 bond and 3 angle
# tables are converted .xvg -> .pot with the reference's converter logic (chemlab_b200.espressopp.tools.convert.gromacs ==
# tools/convert_gromacs2espp.py, byte-exact on the shipped pairs) and packed into one compressed npz (710 KB instead of 2.8 MB of
# text; "%15.8g" text round-trips exactly through float64).  table_a1 / table_a2 are missing from the reference
# (.MISSING_LARGE_BLOBS) and are copies of the shipped table_a3.
def pack_rim135():
    import glob
    import sys
    import tempfile
    import numpy as np
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from chemlab_b200.espressopp.tools.convert.gromacs import convertTable
    src = os.path.join(REF, "examples", "rim135")
    dst = os.path.join(HERE, "rim135")
    os.makedirs(dst, exist_ok=True)
    for f in ("cg_conf.gro", "cg_topol.top", "params", "reaction.cfg"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    arrays = {}
    with tempfile.TemporaryDirectory() as tmp:
        for xvg in sorted(glob.glob(os.path.join(src, "table_*.xvg"))):
            name = os.path.basename(xvg)[:-4]
            pot = os.path.join(tmp, name + ".pot")
            convertTable(xvg, pot)
            arrays[name] = np.loadtxt(pot)
    for k in (1, 2):
        arrays["table_a%d" % k] = arrays["table_a3"]
    np.savez_compressed(os.path.join(dst, "tables.npz"), **arrays)
    print("rim135: %d tables packed" % len(arrays))


# hyperbranched (config 3 base system): inputs as shipped + the shipped .pot tables packed like rim135's.  The angle and dihedral
# tables are missing from the reference (.MISSING_LARGE_BLOBS); the test synthesises smooth ones (tests/test_gpu_driver.py).
def pack_hyperbranched():
    import glob
    import numpy as np
    src = os.path.join(REF, "examples", "hyperbranched")
    dst = os.path.join(HERE, "hyperbranched")
    os.makedirs(dst, exist_ok=True)
    for f in ("conf.gro", "topol.top", "ffnb.itp", "params", "reaction.cfg"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    arrays = {os.path.basename(p)[:-4]: np.loadtxt(p) for p in sorted(glob.glob(os.path.join(src, "table_*.pot")))}
    np.savez_compressed(os.path.join(dst, "tables.npz"), **arrays)
    print("hyperbranched: %d tables packed" % len(arrays))


# dacron (config 4 base system): examples/dacron/no_water/test_1 as shipped (topology with its two .itp includes, coordinates,
# arg-file, reaction config, exclusion list).  Pair tables: the 14 shipped .xvg of test_1 converted with the reference's converter
# logic; the 7 type pairs test_1 lacks (C_D, D_E and the water pairs *_W) are the shipped .pot files of
# examples/dacron/rev_with_water/test_3 (same force field).  Angle tables a0..a3 (45,000 rows each) and bond table b2 as shipped;
# the dihedral tables d0/d1 are missing from the reference (.MISSING_LARGE_BLOBS) and are synthesised by the users of this fixture.
def pack_dacron():
    import glob
    import sys
    import tempfile
    import numpy as np
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from chemlab_b200.espressopp.tools.convert.gromacs import convertTable
    src = os.path.join(REF, "examples", "dacron", "no_water", "test_1")
    alt = os.path.join(REF, "examples", "dacron", "rev_with_water", "test_3")
    dst = os.path.join(HERE, "dacron")
    os.makedirs(dst, exist_ok=True)
    for f in ("conf.gro", "topol.top", "diol_cg.itp", "ter_cg.itp", "params", "reaction.cfg", "exclusion_topol.list"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    arrays = {}
    with tempfile.TemporaryDirectory() as tmp:
        for xvg in sorted(glob.glob(os.path.join(src, "table_*.xvg"))):
            name = os.path.basename(xvg)[:-4]
            pot = os.path.join(tmp, name + ".pot")
            convertTable(xvg, pot)
            arrays[name] = np.loadtxt(pot)
    for pot in sorted(glob.glob(os.path.join(alt, "table_?_?.pot"))):
        name = os.path.basename(pot)[:-4]
        if name not in arrays:
            arrays[name] = np.loadtxt(pot)
    np.savez_compressed(os.path.join(dst, "tables.npz"), **arrays)
    print("dacron: %d tables packed: %s" % (len(arrays), " ".join(sorted(arrays))))


# dacron/restrict: the RestrictReaction example (group key `connectivity_map:connections.list`, reaction_setup.py:74-75,115-126).
# Only the files that differ from no_water/test_1 are stored (coordinates, topology, arg-file, reaction config, the connectivity map);
# the two .itp includes, the exclusion list and every table are byte-identical to the dacron fixture and are taken from it by
# chemlab_b200.synthetic.prepare_example.
def pack_dacron_restrict():
    src = os.path.join(REF, "examples", "dacron", "restrict")
    dst = os.path.join(HERE, "dacron_restrict")
    os.makedirs(dst, exist_ok=True)
    for f in ("conf.gro", "topol.top", "params", "reaction.cfg", "connections.list"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        os.chmod(os.path.join(dst, f), 0o644)
    for f in ("diol_cg.itp", "ter_cg.itp", "exclusion_topol.list"):
        assert open(os.path.join(src, f), "rb").read() == open(os.path.join(HERE, "dacron", f), "rb").read(), f
    print("dacron_restrict: copied")


# mf/espp_cg_1: melamine-formaldehyde network, one bead type, A(0,3) + A(0,3) -> A(1):A(1) with intramolecular: 0 (every bead binds
# up to three others, never inside its own molecule), harmonic reaction bonds; inputs as shipped, table_A_A.xvg converted with
# the reference's converter logic and packed like the other examples.
def pack_mf():
    import sys
    import tempfile
    import numpy as np
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from chemlab_b200.espressopp.tools.convert.gromacs import convertTable
    src = os.path.join(REF, "examples", "mf", "espp_cg_1")
    dst = os.path.join(HERE, "mf")
    os.makedirs(dst, exist_ok=True)
    for f in ("conf.gro", "topol.top", "params", "reaction.cfg"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        os.chmod(os.path.join(dst, f), 0o644)
    with tempfile.TemporaryDirectory() as tmp:
        pot = os.path.join(tmp, "table_A_A.pot")
        convertTable(os.path.join(src, "table_A_A.xvg"), pot)
        np.savez_compressed(os.path.join(dst, "tables.npz"), table_A_A=np.loadtxt(pot))
    print("mf: packed")


# pccg_lj/chemical_reactions: Kremer-Grest-like divinyl monomers in solvent (15,200 beads), pair-specific LJ, FENE + LJ bonds
# ([ bonds ] func 9) also for the reaction bonds, Cosine angles (func 11) generated by the TopologyManager, ATRP activator.
# Inputs as shipped; hooks.py of the fixture is a Python-3 re-authoring (the shipped one is Python 2) and is NOT copied.
def pack_pccg_lj():
    src = os.path.join(REF, "examples", "pccg_lj", "chemical_reactions")
    dst = os.path.join(HERE, "pccg_lj")
    os.makedirs(dst, exist_ok=True)
    for f in ("conf.gro", "topol.top", "solvent.itp", "atrp.cfg", "params", "exclusion_topol.list"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        os.chmod(os.path.join(dst, f), 0o644)
    print("pccg_lj: copied")


if __name__ == "__main__":
    pack_rim135()
    pack_hyperbranched()
    pack_dacron()
    pack_dacron_restrict()
    pack_mf()
    pack_pccg_lj()
