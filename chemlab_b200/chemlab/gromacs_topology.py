"""Topology preparation and interaction assembly: the specification of WHICH kernels exist for a system.

Behaviour follows src/chemlab/gromacs_topology.py: `#include/#define` preprocessing, molecule
replication with 1-based atom ids, type ids in order of first appearance, exclusions from `nrexcl`, GROMACS `func` codes
mapped onto espressopp interaction classes, static (per-parameter) versus dynamic (type-dispatched) tuple lists."""
import collections
import math
import os

from .. import espressopp
from . import files_io
from .py2compat import py2_dict_order  # noqa: F401  (type ids follow the reference's Python-2 dict order)


# ------------------------------------------------------------------------------------------ preprocessing
def preprocess(path, defines=None, cwd="."):
    """Returns the lines of `path` with `#include "x"` expanded (relative to cwd / the include's dirname) and
    `#define NAME value...` collected (gromacs_topology.py:60-85)."""
    defines = {} if defines is None else defines
    lines = []
    with open(os.path.join(cwd, path)) as f:
        for line in f:
            if line.startswith(";"):
                continue
            if "include" in line:
                name = line.split()[1].strip('"')
                inc = os.path.join(cwd, name)
                if not os.path.exists(inc) and os.path.exists(name):
                    inc = name
                inc_lines, _ = preprocess(os.path.basename(inc), defines, os.path.dirname(inc) or ".")
                lines += inc_lines
            elif "define" in line:
                t = line.split()
                if len(t) > 2:
                    defines[t[1]] = " ".join(t[2:])
            elif line.strip():
                lines.append(line.rstrip("\n"))
    return lines, defines


def substitute_defines(lines, defines):
    """Whole-token substitution of #define names (gromacs_topology.py:88-107)."""
    out = []
    for line in lines:
        s = line.strip()
        if not s:
            continue
        if not s.startswith((";", "#")):
            hit = set(s.split()) & set(defines)
            if hit:
                k = hit.pop()
                s = s.replace(k, defines[k])
        out.append(s)
    return out


def convertc6c12(c6, c12, cr):
    """(c6, c12) -> (sigma, epsilon) for combination rule 1 only (gromacs_topology.py:110-121)."""
    if cr != 1:
        return c6, c12
    if c12 == 0.0:
        return 1.0, 0.0
    sig = (c12 / c6) ** (1.0 / 6.0)
    return sig, (0.25 * c6 * sig ** -6.0 if sig > 0.0 else 0.0)


def combination(s1, e1, s2, e2, cr):
    """rule 2: arithmetic sigma; otherwise geometric (gromacs_topology.py:452-460)."""
    return (0.5 * (s1 + s2) if cr == 2 else math.sqrt(s1 * s2)), math.sqrt(e1 * e2)


def pot_table(name):
    """`table_X.xvg` is converted to `.pot` on first use; a shipped `.pot` wins over its `.xvg` (appendix A)."""
    pot = "%s.pot" % name.replace(".xvg", "").replace(".pot", "")
    if not os.path.exists(pot):
        print("Convert %s to %s" % (name, pot))
        espressopp.tools.convert.gromacs.convertTable(name, pot)
    return pot


# ------------------------------------------------------------------------------------------ topology
class GromacsTopology:
    def __init__(self, input_topol, generate_exclusions=True):
        self.input_file = input_topol
        self.generate_exclusions = generate_exclusions
        self.atoms, self.bonds, self.angles, self.dihedrals, self.pairs = {}, {}, {}, {}, {}
        self.bondparams, self.angleparams, self.dihedralparams = {}, {}, {}
        self.atomsym_atomtype, self.atomtype_atomsym, self.used_atomsym_atomtype = {}, {}, {}
        self.exclusions = set()

    def read(self):
        cwd = os.path.dirname(self.input_file) or "."
        lines, defines = preprocess(os.path.basename(self.input_file), cwd=cwd)
        self.gt = self.topol = files_io.GROMACSTopologyFile(self.input_file)
        self.gt.content = substitute_defines(lines, defines)
        self.gt.read()
        self.master_topol = files_io.GROMACSTopologyFile(self.input_file).read()   # un-preprocessed master file
        self._prepare()
        return self

    def add_new_atomtype(self, atype_id, atype_name, is_used=False):
        self.atomtype_atomsym[atype_id] = atype_name
        self.atomsym_atomtype[atype_name] = atype_id
        if is_used:
            self.used_atomsym_atomtype[atype_name] = atype_id

    def _prepare(self):
        gt = self.gt
        cr = gt.defaults.get("combinationrule", 1)
        self.used_atomnr = set()
        self.used_atomnr2atom_type = collections.defaultdict(set)
        next_type = 0
        offset = 0
        for mol, n_mols in gt.molecules:
            matoms = gt.molecules_data[mol].get("atoms", {})
            per_atom = {}
            for aid in sorted(matoms):
                a = matoms[aid]
                t = gt.atomtypes[a.atom_type]
                if a.atom_type not in self.atomsym_atomtype:           # type ids in order of first appearance
                    self.atomsym_atomtype[a.atom_type] = next_type
                    next_type += 1
                nr = gt.atom_name2atomnr[a.atom_type]
                self.used_atomnr.add(nr)
                self.used_atomnr2atom_type[nr].add(a.atom_type)
                self.used_atomsym_atomtype[a.atom_type] = self.atomsym_atomtype[a.atom_type]
                sig, eps = convertc6c12(t["sigma"], t["epsilon"], cr)
                per_atom[aid] = {"molecule": a.chain_name, "type": a.atom_type, "sig": sig, "eps": eps,
                                 "type_id": self.atomsym_atomtype[a.atom_type], "state": t.get("state", 0),
                                 "charge": a.charge if a.charge else t["charge"], "mass": a.mass if a.mass else t["mass"],
                                 "molecule_name": a.molecule_name, "name": a.name, "cgnr": a.cgnr,
                                 "chain_idx": a.chain_idx, "chain_name": a.chain_name}
            n_atoms = len(matoms)
            for m in range(n_mols):
                for k, v in per_atom.items():
                    self.atoms[offset + k + m * n_atoms] = v
            # molecule by molecule, like `_replicate_lists` (:431-446); within the fixed lists the tuples then follow the
            # insertion order of these dicts (the Python-2 reference walks them in the hash order of the id tuples, which no
            # shipped artefact records)
            for name, store in (("bonds", self.bonds), ("angles", self.angles), ("dihedrals", self.dihedrals), ("pairs", self.pairs)):
                for m in range(n_mols):
                    for tup, params in gt.molecules_data[mol].get(name, {}).items():
                        store[tuple(offset + x + m * n_atoms for x in tup)] = params
            offset += n_mols * n_atoms
        for v in gt.nonbond_params.values():
            if v["func"] == 1 and cr == 1 and v["params"]:
                v["params"][0], v["params"][1] = convertc6c12(float(v["params"][0]), float(v["params"][1]), cr)
        # remaining [atomtypes] of the master file get ids too (reaction products may not appear in any molecule).  The reference
        # walks `master_topol.atomtypes.items()` (:260), a Python-2 dict: the ids follow CPython 2.7's hash order of the type
        # names, reproduced by py2_dict_order -- this gives the ids of the shipped run log examples/atrp_lj/single:199-205
        # (MA0 ML1 DA2 FA3 PA4 RA5 PL6 for the file order MA ML PA FA DA RA PL; tests/test_runlog_cpu.py)
        master = self.master_topol.atom_name2atomnr
        all_types = {name: master[name] for name in py2_dict_order(list(master))}
        for name, nr in gt.atom_name2atomnr.items():      # also the types that arrive through #include files
            all_types.setdefault(name, nr)
        for name, nr in all_types.items():
            self.used_atomnr.add(nr)
            self.used_atomnr2atom_type[nr].add(name)
            if name not in self.atomsym_atomtype:
                self.atomsym_atomtype[name] = next_type
                next_type += 1
            self.used_atomsym_atomtype[name] = self.atomsym_atomtype[name]
        self.atomtype_atomsym = {v: k for k, v in self.atomsym_atomtype.items()}
        self._prepare_bondedparams()
        if self.generate_exclusions:
            self._prepare_exclusions()

    def _type_ids(self, nr):
        return [self.atomsym_atomtype[t] for t in self.used_atomnr2atom_type[nr]]

    def _prepare_bondedparams(self):
        """[bondtypes]/[angletypes]/[dihedraltypes] keyed by canonical type-id tuples for the type-dispatched lists."""
        gt, used = self.gt, self.used_atomnr
        for i, row in gt.bondtypes.items():
            for j, p in row.items():
                if i in used and j in used:
                    for t1 in self._type_ids(i):
                        for t2 in self._type_ids(j):
                            self.bondparams[tuple(sorted((t1, t2)))] = p
        for i, r1 in gt.angletypes.items():
            for j, r2 in r1.items():
                for k, p in r2.items():
                    if i in used and j in used and k in used:
                        for t1 in self._type_ids(i):
                            for t2 in self._type_ids(j):
                                for t3 in self._type_ids(k):
                                    self.angleparams[(t3, t2, t1) if t1 > t3 else (t1, t2, t3)] = p
        for i, r1 in gt.dihedraltypes.items():
            for j, r2 in r1.items():
                for k, r3 in r2.items():
                    for l, p in r3.items():
                        if {i, j, k, l} <= used:
                            for t1 in self._type_ids(i):
                                for t2 in self._type_ids(j):
                                    for t3 in self._type_ids(k):
                                        for t4 in self._type_ids(l):
                                            self.dihedralparams[(t4, t3, t2, t1) if t4 > t1 else (t1, t2, t3, t4)] = p

    @staticmethod
    def molecule_exclusions(bonds, nrexcl):
        """Pairs at most `nrexcl` bonds apart inside one molecule (gromacs_topology.py:317-377)."""
        adj = collections.defaultdict(set)
        for a, b in bonds:
            adj[a].add(b); adj[b].add(a)
        out = {tuple(sorted(b)) for b in bonds}
        for root in adj:
            seen, frontier = {root}, {root}
            for _ in range(nrexcl):
                frontier = {y for x in frontier for y in adj[x]} - seen
                seen |= frontier
            out |= {tuple(sorted((root, x))) for x in seen if x != root}
        return out

    def _prepare_exclusions(self):
        self.exclusions = {tuple(sorted(b)) for b in self.bonds}
        offset = 0
        for mol, n_mols in self.gt.molecules:
            n_atoms = len(self.gt.molecules_data[mol].get("atoms", {}))
            mb = self.gt.molecules_data[mol].get("bonds")
            if mb:
                for pair in self.molecule_exclusions(list(mb), self.gt.moleculetype[mol]):
                    for m in range(n_mols):
                        self.exclusions.add(tuple(sorted(offset + x + m * n_atoms for x in pair)))
            # the reference advances this offset by n_mols only (gromacs_topology.py:314), which breaks the second
            # molecule type; shipped multi-molecule inputs side-step it with exclusion_list= files.  Fixed here.
            offset += n_mols * n_atoms


def gen_particle_list(coordinate, topol):
    """(property names, rows) for storage.addParticles (gromacs_topology.py:1418-1441)."""
    props = ["id", "type", "pos", "mass", "q", "res_id", "state", "lambda_adr"]
    rows = []
    for aid in sorted(coordinate.atoms):
        c, t = coordinate.atoms[aid], topol.atoms[aid]
        rows.append([aid, t["type_id"], espressopp.Real3D(*c.position), t["mass"], t["charge"], c.chain_idx, t.get("state", 0), 1.0])
    return props, rows


# ------------------------------------------------------------------------------------------ non-bonded
def set_nonbonded_interactions(system, gt, vl, lj_cutoff, qq_cutoff=None, tab_cutoff=None, tables=None, cr_observs=None):
    """func 1 -> LJ, func 8 / --table_groups -> Tabulated, func 10/12 -> MixedTabulated (doc/topology.rst:137-207;
    gromacs_topology.py:463-899).  Capped/multi/scaled variants (9, 13, 16, 17, 18) are outside the engine's scope."""
    defaults, atomparams = gt.gt.defaults, gt.gt.atomtypes
    sym2id = gt.used_atomsym_atomtype
    tab_cutoff = lj_cutoff if tab_cutoff is None else tab_cutoff
    tables = tables.split(",") if isinstance(tables, str) else (tables or [])
    cr_observs = {} if cr_observs is None else cr_observs
    cr = int(defaults.get("combinationrule", 1))
    lj = espressopp.interaction.VerletListLennardJones(vl)
    tab = espressopp.interaction.VerletListTabulated(vl)
    mixed = espressopp.interaction.VerletListMixedTabulated(vl)
    used = {"lj": False, "tab": False, "mixed": False}
    pairs = sorted({tuple(sorted((a, b))) for a in sym2id for b in sym2id})
    print("Number of non-bonded type pairs: %d" % len(pairs))
    for n1, n2 in pairs:
        t1, t2 = sym2id[n1], sym2id[n2]
        param = gt.gt.nonbond_params.get((n1, n2))
        table_name, sig, eps = None, -1.0, -1.0
        if param:
            func, pr = param["func"], param["params"]
            if func == 1:
                if pr:
                    sig, eps = float(pr[0]), float(pr[1])
                else:
                    sig, eps = combination(atomparams[n1]["sigma"], atomparams[n1]["epsilon"], atomparams[n2]["sigma"], atomparams[n2]["epsilon"], cr)
            elif func == 8:
                table_name = pr[0] if pr else "table_%s_%s.xvg" % (n1, n2)
            elif func == 10:     # tab1 tab2 type total : U = x tab1 + (1-x) tab2, x = N(type)/total
                cr_type, cr_total = sym2id[pr[2]], int(pr[3])
                key = (cr_type, cr_total, None)
                if key not in cr_observs:
                    cr_observs[key] = espressopp.analysis.ChemicalConversion(system, cr_type, cr_total)
                print("Set mixed tabulated potential %s-%s with conversion observable (U=x*%s + (1-x)*%s)" % (t1, t2, pot_table(pr[0]), pot_table(pr[1])))
                mixed.setPotential(type1=t1, type2=t2, potential=espressopp.interaction.MixedTabulated(
                    itype=1, tab1=pot_table(pr[0]), tab2=pot_table(pr[1]), cr_observation=cr_observs[key], cutoff=tab_cutoff))
                used["mixed"] = True
                continue
            elif func == 12:     # tab1 tab2 mix_value
                print("Set mixed tabulated potential %s-%s with static scaling x=%s (U=%s*%s+(1-%s)*%s)" % (t1, t2, pr[2], pr[2], pot_table(pr[0]), pr[2], pot_table(pr[1])))
                mixed.setPotential(type1=t1, type2=t2, potential=espressopp.interaction.MixedTabulated(
                    itype=1, tab1=pot_table(pr[0]), tab2=pot_table(pr[1]), mix_value=float(pr[2]), cutoff=tab_cutoff))
                used["mixed"] = True
                continue
            elif func in (9, 13, 16, 17, 18):
                raise NotImplementedError("nonbond_params func %d (multi/capped/scaled tables) is outside the scope of the B200 engine" % func)
            elif func in (11, 14, 15):
                raise NotImplementedError("nonbond_params func %d (dynamic-resolution / pair-scaled potentials of the AdResS-style machinery) is "
                                          "outside the scope of the B200 engine (SURVEY E21)" % func)
            else:
                raise RuntimeError("Functional %d not found" % func)
        elif n1 in tables and n2 in tables:
            table_name = "table_%s_%s.xvg" % (n1, n2)
        else:
            sig, eps = combination(atomparams[n1]["sigma"], atomparams[n1]["epsilon"], atomparams[n2]["sigma"], atomparams[n2]["epsilon"], cr)
        if table_name is not None:
            print("Set tab potential %s-%s: %s" % (n1, n2, table_name))
            tab.setPotential(type1=t1, type2=t2, potential=espressopp.interaction.Tabulated(itype=1, filename=pot_table(table_name), cutoff=tab_cutoff))
            used["tab"] = True
        elif sig > 0.0:
            print("Set LJ potential %s-%s, eps=%s, sig=%s, cutoff=%s" % (n1, n2, eps, sig, lj_cutoff))
            lj.setPotential(type1=t1, type2=t2, potential=espressopp.interaction.LennardJones(epsilon=eps, sigma=sig, cutoff=lj_cutoff))
            used["lj"] = True
    # registration order of the reference: lj-mix_tab (:757-790), coulomb (:866-878), lj, lj-tab (:883-893)
    if used["mixed"]:
        system.addInteraction(mixed, "lj-mix_tab")
    # the `coulomb` term: registered whenever the cutoff is positive, zero for the (neutral) coarse-grained beads
    fudge_qq = float(defaults.get("fudgeQQ", 1.0))
    if qq_cutoff is not None and float(qq_cutoff) > 0.0 and 138.935485 * fudge_qq > 0.0:
        pot_qq = espressopp.interaction.CoulombTruncated(prefactor=138.935485 * fudge_qq, cutoff=float(qq_cutoff))
        coul = espressopp.interaction.VerletListCoulombTruncated(vl)
        for n1, n2 in pairs:
            coul.setPotential(type1=sym2id[n1], type2=sym2id[n2], potential=pot_qq)
        system.addInteraction(coul, "coulomb")
    if used["lj"]:
        system.addInteraction(lj, "lj")
    if used["tab"]:
        system.addInteraction(tab, "lj-tab")
    return cr_observs, []


# ------------------------------------------------------------------------------------------ bonded
def _deg(x):
    return float(x) * math.pi / 180.0


def _bond_pot(func, raw):
    I = espressopp.interaction
    if func == 1:
        return I.Harmonic(K=float(raw[1]) / 2.0, r0=float(raw[0]))              # chemlab halves the GROMACS K (:918)
    if func == 8:
        return I.Tabulated(itype=1, filename=pot_table("table_b%d.xvg" % int(float(raw[0]))))
    if func == 7:
        return I.FENE(K=float(raw[1]), r0=0.0, rMax=float(raw[0]))
    if func == 9:                                                                # FENE + LJ: rMax K sigma epsilon (:935-944)
        return I.FENELennardJones(K=float(raw[1]), r0=0.0, rMax=float(raw[0]), sigma=float(raw[2]), epsilon=float(raw[3]))
    raise RuntimeError("Unknown bond func type %s" % func)


def _angle_pot(func, raw):
    I = espressopp.interaction
    if func == 1:
        return I.AngularHarmonic(K=float(raw[1]) / 2.0, theta0=_deg(raw[0]))    # :1073
    if func == 8:
        return I.TabulatedAngular(itype=1, filename=pot_table("table_a%d.xvg" % int(float(raw[0]))))
    if func == 11:
        return I.Cosine(K=float(raw[1]), theta0=_deg(raw[0]))                   # :1082, K un-halved
    raise RuntimeError("Unknown angle func type %s" % func)


def _dihedral_pot(func, raw):
    I = espressopp.interaction
    if func == 8:
        return I.TabulatedDihedral(itype=1, filename=pot_table("table_d%d.xvg" % int(float(raw[0]))))
    if func == 12:
        return I.DihedralHarmonic(K=float(raw[1]), phi0=_deg(raw[0]))
    raise NotImplementedError("dihedral func %s (RB / n-cos) is outside the scope of the B200 engine" % func)


_KINDS = {
    2: dict(params="bondparams", tuples="bonds", make=_bond_pot, List=lambda: espressopp.FixedPairList, add="addBonds", label="bonds",
            static={1: "FixedPairListHarmonic", 7: "FixedPairListFENE", 8: "FixedPairListTabulated", 9: "FixedPairListFENELennardJones"},
            typed={1: "FixedPairListTypesHarmonic", 7: "FixedPairListTypesFENE", 8: "FixedPairListTypesTabulated", 9: "FixedPairListTypesFENELennardJones"}),
    3: dict(params="angleparams", tuples="angles", make=_angle_pot, List=lambda: espressopp.FixedTripleList, add="addTriples", label="angles",
            static={1: "FixedTripleListAngularHarmonic", 8: "FixedTripleListTabulatedAngular", 11: "FixedTripleListCosine"},
            typed={1: "FixedTripleListTypesAngularHarmonic", 8: "FixedTripleListTypesTabulatedAngular", 11: "FixedTripleListTypesCosine"}),
    4: dict(params="dihedralparams", tuples="dihedrals", make=_dihedral_pot, List=lambda: espressopp.FixedQuadrupleList, add="addQuadruples",
            label="dihedrals", static={8: "FixedQuadrupleListTabulatedDihedral", 12: "FixedQuadrupleListDihedralHarmonic"},
            typed={8: "FixedQuadrupleListTypesTabulatedDihedral", 12: "FixedQuadrupleListTypesDihedralHarmonic"}),
}

DynKey = collections.namedtuple("DynKey", "func is_observe_list")


def _canon(t):
    t = tuple(t)
    return t if t <= t[::-1] else t[::-1]


def _set_tuple_interactions(arity, system, gt, dynamic_type_ids, change_types, separate=frozenset(), name=None):
    """A tuple whose type tuple involves a DYNAMIC type (reactant / product / neighbour-changed type) goes into a
    type-dispatched FixedXListTypes* interaction; the others into static per-parameter lists (appendix A;
    gromacs_topology.py:969-1011,1102-1135,1230-1264)."""
    K = _KINDS[arity]
    name = name or K["label"]
    params = getattr(gt, K["params"])
    dyn_types = {}
    by_func = collections.defaultdict(list)
    # The reference's rule, kept to the letter (:969-1011, :1102-1135, :1230-1264): the parameter dicts are keyed by type tuples in
    # ONE orientation (bonds sorted, angles first <= last, dihedrals first >= last: _prepare_bondedparams), a bond is looked up by
    # its sorted types, but an angle / dihedral by its types in the order the tuple is LISTED -- a tuple listed in the other
    # orientation is therefore static (with the parameters of its line, or those of the reversed key) even when its types are
    # dynamic.  examples/hyperbranched: the 1000 listed dihedrals `2 1 3 4 8 0.0 1.0` stay on table d0 whatever their types become.
    for pt, p in params.items():
        if not (set(pt) & set(dynamic_type_ids)) and tuple(pt) not in change_types:
            continue
        dyn_types[tuple(sorted(pt)) if arity == 2 else tuple(pt)] = p
        by_func[p["func"]].append((pt, p))
    dyn_tuples = collections.defaultdict(list)
    static = collections.defaultdict(lambda: collections.defaultdict(list))
    for tup, raw in getattr(gt, K["tuples"]).items():
        pt = tuple(gt.atoms[x]["type_id"] for x in tup)
        if arity == 2:
            pt = tuple(sorted(pt))
        if raw:
            func, pr = int(raw[0]), tuple(float(x) for x in raw[1:])
        else:
            p = params.get(pt) or params.get(pt[::-1])
            if p is None:
                raise RuntimeError("no %s parameters for types %s" % (K["label"], pt))
            func, pr = int(p["func"]), tuple(float(x) for x in p["params"])
        if pt in dyn_types and pt not in separate:
            dyn_tuples[func].append(tup)
        else:
            static[func][pr].append(tup)
    for func in by_func:
        dyn_tuples.setdefault(func, [])
    count = 0
    dynamic_lists, static_lists = {}, []
    I = espressopp.interaction
    for func, tuples in dyn_tuples.items():
        lst = K["List"]()(system.storage)
        getattr(lst, K["add"])(tuples)
        inter = getattr(I, K["typed"][func])(system, lst)
        observe = False
        lst.params = collections.defaultdict(dict)
        for pt, p in by_func.get(func, []):
            observe = observe or tuple(pt) in change_types
            kw = {"type%d" % (i + 1): t for i, t in enumerate(pt)}
            inter.setPotential(potential=K["make"](func, p["params"]), **kw)
        system.addInteraction(inter, "dyn_%s_%d" % (name, count))
        count += 1
        dynamic_lists[DynKey(func, observe)] = lst
    for func, groups in static.items():
        for pr, tuples in groups.items():
            lst = K["List"]()(system.storage)
            getattr(lst, K["add"])(tuples)
            lst.params = (func, pr)
            system.addInteraction(getattr(I, K["static"][func])(system, lst, K["make"](func, pr)), "%s_%d" % (name, count))
            count += 1
            static_lists.append(lst)
    return dynamic_lists, static_lists


def set_bonded_interactions(system, gt, dynamic_type_ids, change_bond_types=frozenset(), separate_fpls=frozenset(), name="bonds"):
    d, s = _set_tuple_interactions(2, system, gt, dynamic_type_ids, change_bond_types, separate_fpls, name)
    return d, s, []


def set_angle_interactions(system, gt, dynamic_type_ids, change_angle_types=frozenset(), name="angles"):
    return _set_tuple_interactions(3, system, gt, dynamic_type_ids, change_angle_types, name=name)


def set_dihedral_interactions(system, gt, dynamic_type_ids, change_dihedral_types=frozenset(), name="dihedrals"):
    return _set_tuple_interactions(4, system, gt, dynamic_type_ids, change_dihedral_types, name=name)


def set_pair_interactions(system, gt, args, dynamic_type_ids):
    """1-4 `[ pairs ]` as Lennard-Jones on pair lists (gromacs_topology.py:1314-1411): pairs whose types can change in a reaction
    go to ONE FixedPairListTypesLennardJones (`dyn_lj14_<k>`: sigma/epsilon by the combination rule, epsilon scaled by fudgeLJ);
    the other pairs are grouped by their parameters into FixedPairListLennardJones interactions `lj14_<k>` -- (sigma, epsilon)
    written on the pair line, else the combination rule with fudgeLJ when `gen-pairs` is set.  (The reference calls
    `combination` with three arguments on that static branch, :1358, and would stop there; the parameters are used as given
    instead.)  The Coulomb part of the 1-4 term needs charged particles and stays outside the scope (all shipped CG beads are neutral)."""
    if not gt.pairs:
        return None, []
    defaults, atomparams = gt.gt.defaults, gt.gt.atomtypes
    cr = int(defaults.get("combinationrule", 1))
    fudge = float(defaults.get("fudgeLJ", 1.0))
    cutoff = float(args.lj_cutoff)
    id2sym = {v: k for k, v in gt.atomsym_atomtype.items()}
    static, dynamic = {}, []
    for b, parameters in gt.pairs.items():
        ptypes = [gt.atoms[x]["type_id"] for x in b]
        if set(ptypes) & set(dynamic_type_ids):
            dynamic.append(b)
            continue
        params = tuple(float(x) for x in parameters[1:])
        if not params:
            if not defaults.get("gen-pairs"):
                raise RuntimeError("pair %s has no parameters and [ defaults ] gen-pairs is off" % (b,))
            n1, n2 = id2sym[ptypes[0]], id2sym[ptypes[1]]
            sig, eps = combination(atomparams[n1]["sigma"], atomparams[n1]["epsilon"], atomparams[n2]["sigma"], atomparams[n2]["epsilon"], cr)
            params = (sig, fudge * eps)
        static.setdefault(params[:2], []).append(b)
    static_fpls, count = [], 0
    for (sig, eps), b_list in sorted(static.items()):
        fpl = espressopp.FixedPairList(system.storage)
        fpl.addBonds(b_list)
        inter = espressopp.interaction.FixedPairListLennardJones(system, fpl, espressopp.interaction.LennardJones(epsilon=eps, sigma=sig, cutoff=cutoff))
        system.addInteraction(inter, "lj14_%d" % count)
        static_fpls.append(fpl)
        count += 1
    dfpl = espressopp.FixedPairList(system.storage)
    if dynamic:
        dfpl.addBonds(dynamic)
    type_pairs = sorted({tuple(sorted((a, b))) for a in gt.used_atomsym_atomtype for b in gt.used_atomsym_atomtype
                         if gt.used_atomsym_atomtype[a] in dynamic_type_ids or gt.used_atomsym_atomtype[b] in dynamic_type_ids})
    if type_pairs:
        print("Set up 1-4 pair interactions")
        inter = espressopp.interaction.FixedPairListTypesLennardJones(system, dfpl)
        for n1, n2 in type_pairs:
            sig, eps = combination(atomparams[n1]["sigma"], atomparams[n1]["epsilon"], atomparams[n2]["sigma"], atomparams[n2]["epsilon"], cr)
            inter.setPotential(type1=gt.used_atomsym_atomtype[n1], type2=gt.used_atomsym_atomtype[n2],
                               potential=espressopp.interaction.LennardJones(sigma=sig, epsilon=fudge * eps, cutoff=cutoff))
        system.addInteraction(inter, "dyn_lj14_%d" % count)
    return dfpl, static_fpls


def set_coulomb_interactions(system, gt, args):
    return None
