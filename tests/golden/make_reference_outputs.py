"""Outputs of the REFERENCE's own Python code for the host-side formats, generated here (the reference tree is mounted in this
container only) and committed as tests/golden/refout/reference_outputs.json: the prompt's rule for a Python reference -- import
it where it can be imported, commit the vectors together with the script that made them.

Four modules of /root/reference/src run under Python 3 once the Python-2 builtins they rely on are supplied (list-returning
map / filter / zip, xrange, the ConfigParser module name): chemlab/reaction_parser.py, chemlab/files_io.py, app_args.py.  For every
fixture under tests/golden/ that is a copy of a shipped input, the reference code is run on the ORIGINAL file and its result stored:
  reaction configs -> parse_config(...)                       (reaction_parser.py:235-266)
  arg files        -> vars(_args().parse_args(['@params']))    (app_args.py:71-211; rng_seed dropped: its default is random)
  topologies       -> GROMACSTopologyFile(...).read()          (files_io.py:401-821): every parsed section
  coordinates      -> GROFile(...).read() digest, and the bytes GROFile.write produces for the same content (files_io.py:158-257)
Nothing here is imported by the product or by a test; tests/test_reference_outputs_cpu.py reads the JSON."""
import builtins
import configparser
import contextlib
import hashlib
import importlib.util
import io
import json
import os
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
TOPO_FIELDS = ("atom_id", "atom_type", "chain_idx", "chain_name", "name", "cgnr", "charge", "mass", "molecule_name")
TOPO_SECTIONS = ("defaults", "atomtypes", "bondtypes", "angletypes", "dihedraltypes", "nonbond_params", "atomstate", "moleculetype", "molecules",
                 "system_name", "atom_name2atomnr")
# fixture (relative to tests/golden) -> original (relative to /root/reference)
CFG = {"atrp_lj/atrp.cfg": "examples/atrp_lj/atrp.cfg", "chain_growth_catalytic/reaction.cfg": "examples/chain_growth_catalytic/reaction.cfg",
       "rim135/reaction.cfg": "examples/rim135/reaction.cfg", "hyperbranched/reaction.cfg": "examples/hyperbranched/reaction.cfg",
       "dacron/reaction.cfg": "examples/dacron/no_water/test_1/reaction.cfg", "dacron_restrict/reaction.cfg": "examples/dacron/restrict/reaction.cfg",
       "mf/reaction.cfg": "examples/mf/espp_cg_1/reaction.cfg", "pccg_lj/atrp.cfg": "examples/pccg_lj/chemical_reactions/atrp.cfg"}
PARAMS = {"atrp_lj/params": "examples/atrp_lj/params", "chain_growth_catalytic/params": "examples/chain_growth_catalytic/params",
          "rim135/params": "examples/rim135/params", "hyperbranched/params": "examples/hyperbranched/params",
          "dacron/params": "examples/dacron/no_water/test_1/params", "dacron_restrict/params": "examples/dacron/restrict/params",
          "mf/params": "examples/mf/espp_cg_1/params", "pccg_lj/params": "examples/pccg_lj/chemical_reactions/params"}
TOPS = {"atrp_lj/topol.top": "examples/atrp_lj/topol.top", "chain_growth_catalytic/topol.top": "examples/chain_growth_catalytic/topol.top",
        "rim135/cg_topol.top": "examples/rim135/cg_topol.top", "hyperbranched/topol.top": "examples/hyperbranched/topol.top",
        "dacron/topol.top": "examples/dacron/no_water/test_1/topol.top", "dacron_restrict/topol.top": "examples/dacron/restrict/topol.top",
        "mf/topol.top": "examples/mf/espp_cg_1/topol.top", "pccg_lj/topol.top": "examples/pccg_lj/chemical_reactions/topol.top",
        "parser/topol.top": "src/tests/topol.top"}
GROS = {"atrp_lj/conf.gro": "examples/atrp_lj/conf.gro", "chain_growth_catalytic/conf.gro": "examples/chain_growth_catalytic/conf.gro",
        "rim135/cg_conf.gro": "examples/rim135/cg_conf.gro", "hyperbranched/conf.gro": "examples/hyperbranched/conf.gro",
        "dacron/conf.gro": "examples/dacron/no_water/test_1/conf.gro", "dacron_restrict/conf.gro": "examples/dacron/restrict/conf.gro",
        "mf/conf.gro": "examples/mf/espp_cg_1/conf.gro", "pccg_lj/conf.gro": "examples/pccg_lj/chemical_reactions/conf.gro"}


def load_ref(name, rel):
    sys.modules.setdefault("ConfigParser", configparser)
    if not hasattr(configparser, "SafeConfigParser"):
        configparser.SafeConfigParser = configparser.ConfigParser
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    m.map = lambda f, *a: list(builtins.map(f, *a))
    m.filter = lambda f, a: list(builtins.filter(f, a))
    m.zip = lambda *a: list(builtins.zip(*a))
    m.xrange = range
    spec.loader.exec_module(m)
    return m


def norm(o):
    """JSON-able, order-free view of the parsed structures (dict keys as strings, sets sorted, TopoAtom by its fields)."""
    if isinstance(o, dict):
        return {(" ".join(str(x) for x in k) if isinstance(k, tuple) else str(k)): norm(v) for k, v in o.items()}
    if isinstance(o, (set, frozenset)):
        return sorted(norm(v) for v in o)
    if isinstance(o, (list, tuple)):
        return [norm(v) for v in o]
    if type(o).__name__ == "TopoAtom":
        return {n: getattr(o, n, None) for n in TOPO_FIELDS}
    if isinstance(o, float) or isinstance(o, int) or isinstance(o, str) or o is None or isinstance(o, bool):
        return o
    return str(o)


def gro_digest(g):
    rows = []
    for k in sorted(g.atoms):
        a = g.atoms[k]
        v = None if (a.velocity is None or a.velocity[0] is None) else [float(x) for x in a.velocity]
        rows.append([int(a.atom_id), a.name, a.chain_name, int(a.chain_idx), [float(x) for x in a.position], v])
    blob = json.dumps(rows, sort_keys=True).encode()
    return {"n": len(rows), "title": g.title, "box": [float(x) for x in g.box], "first": rows[0], "last": rows[-1], "sha1": hashlib.sha1(blob).hexdigest()}


def topology_view(t):
    out = {s: norm(getattr(t, s, None)) for s in TOPO_SECTIONS}
    if not out["defaults"]:
        out["defaults"] = None          # a master file without [ defaults ]: None in the reference, {} here
    out["molecules_data"] = {mol: {sec: norm(rows) for sec, rows in data.items() if rows} for mol, data in t.molecules_data.items()}
    return out


def main():
    rp = load_ref("ref_reaction_parser", "src/chemlab/reaction_parser.py")
    fio = load_ref("ref_files_io", "src/chemlab/files_io.py")
    aa = load_ref("ref_app_args", "src/app_args.py")
    out = {"cfg": {}, "params": {}, "top": {}, "gro": {}}
    sink = io.StringIO()
    cwd = os.getcwd()
    for fix, rel in CFG.items():
        assert open(os.path.join(HERE, fix), "rb").read() == open(os.path.join(REF, rel), "rb").read(), fix
        with contextlib.redirect_stdout(sink):
            out["cfg"][fix] = norm(rp.parse_config(os.path.join(REF, rel)))
    for fix, rel in PARAMS.items():
        assert open(os.path.join(HERE, fix), "rb").read() == open(os.path.join(REF, rel), "rb").read(), fix
        os.chdir(os.path.dirname(os.path.join(REF, rel)))
        v = vars(aa._args().parse_args(["@params"]))
        v.pop("rng_seed", None)
        out["params"][fix] = norm(v)
    for fix, rel in TOPS.items():
        assert open(os.path.join(HERE, fix), "rb").read() == open(os.path.join(REF, rel), "rb").read(), fix
        os.chdir(os.path.dirname(os.path.join(REF, rel)))
        with contextlib.redirect_stdout(sink):
            t = fio.GROMACSTopologyFile(os.path.basename(rel)); t.read()
        out["top"][fix] = topology_view(t)
    os.chdir(cwd)
    for fix, rel in GROS.items():
        assert open(os.path.join(HERE, fix), "rb").read() == open(os.path.join(REF, rel), "rb").read(), fix
        g = fio.GROFile(os.path.join(REF, rel)); g.read()
        d = gro_digest(g)
        with tempfile.TemporaryDirectory() as tmp:
            g.write(os.path.join(tmp, "w.gro"), force=True)
            d["written_sha1"] = hashlib.sha1(open(os.path.join(tmp, "w.gro"), "rb").read()).hexdigest()
        out["gro"][fix] = d
    path = os.path.join(HERE, "refout", "reference_outputs.json")
    with open(path, "w") as f:
        json.dump(out, f, sort_keys=True, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
