"""`.gro` coordinate files and GROMACS-like `.top/.itp` topology files (formats: SURVEY appendix A;
reference behaviour: src/chemlab/files_io.py:158-308 GROFile, :401-976 GROMACSTopologyFile)."""
import collections
import re

import numpy as np

Atom = collections.namedtuple("Atom", "atom_id name chain_name chain_idx position velocity")


class GROFile:
    """Fixed-column .gro: resid[0:5] resname[5:10] atom[10:15] id[15:20] x[20:28] y[28:36] z[36:44] (+ velocities);
    last line = box.  Particle ids are the file's ids."""
    def __init__(self, file_name):
        self.file_name = file_name
        self.title = ""
        self.atoms = {}
        self.box = None

    def read(self):
        with open(self.file_name) as f:
            lines = f.readlines()
        self.title = lines[0].rstrip("\r\n")
        n = int(lines[1])
        for line in lines[2:n + 2]:
            vel = None
            if len(line) > 45:
                vel = np.array([float(line[44:52]), float(line[52:60]), float(line[60:68])])
            aid = int(line[15:20])
            self.atoms[aid] = Atom(aid, line[10:15].strip(), line[5:10].strip(), int(line[0:5]),
                                   np.array([float(line[20:28]), float(line[28:36]), float(line[36:44])]), vel)
        self.box = np.array([float(x) for x in lines[n + 2].split()][:3])
        return self.atoms

    def update_position(self, system, unfolded=False):
        """Current coordinates out of the engine, with the reference's semantics (files_io.py:261-279): unfolded=True stores
        pos + imageBox * L AND the velocities; unfolded=False replaces the positions only (velocities stay what they were)."""
        ctx = system._ctx
        g = ctx.require_engine().get_particles(fields=("pos", "image", "vel"))
        pids = sorted(ctx.pid)
        box = np.asarray(ctx.box)
        for k, pid in enumerate(pids):
            if pid in self.atoms:
                if unfolded:
                    self.atoms[pid] = self.atoms[pid]._replace(position=g["pos"][k] + g["image"][k] * box, velocity=g["vel"][k])
                else:
                    self.atoms[pid] = self.atoms[pid]._replace(position=g["pos"][k])

    def write(self, file_name=None, force=False, with_velocity=False):
        # the reference's own writer (files_io.py:216-257): title (default 'XXX of molecules'), '%d' atoms, fixed-width rows with
        # '%8.3f' positions AND velocities, box as '%f %f %f'.  The shipped conf.gro of dacron, hyperbranched, rim135,
        # chain_growth_catalytic and pccg_lj were written by it and round-trip byte for byte (tests/test_golden_cpu.py).
        # Ids above 99999 wrap like in GROMACS instead of widening the column.
        out = ["%s\n" % (self.title if self.title else "XXX of molecules"), "%d\n" % len(self.atoms)]
        for aid in sorted(self.atoms):
            a = self.atoms[aid]
            row = "%5d%-5s%5s%5d%8.3f%8.3f%8.3f" % (a.chain_idx % 100000, a.chain_name[:5], a.name[:5], aid % 100000, *a.position)
            if with_velocity and a.velocity is not None:
                row += "%8.3f%8.3f%8.3f" % tuple(a.velocity)
            out.append(row + "\n")
        out.append("%f %f %f\n" % tuple(self.box))
        with open(file_name or self.file_name, "w") as f:
            f.writelines(out)


class TopoAtom:
    __slots__ = ("atom_id", "atom_type", "chain_idx", "chain_name", "name", "cgnr", "charge", "mass", "molecule_name")

    def __init__(self):
        self.charge = None
        self.mass = None


def _nested():
    return collections.defaultdict(_nested)


class GROMACSTopologyFile:
    """Section-driven reader.  Parsed sections: defaults, atomtypes, atomstate, nonbond_params, bondtypes, angletypes,
    dihedraltypes, moleculetype, atoms, bonds, angles, dihedrals (a second consecutive [dihedrals] = impropers), pairs,
    system, molecules.  Unknown sections (e.g. [exclusions]) are skipped like in the reference."""
    def __init__(self, file_name):
        self.file_name = file_name
        self.content = None
        self.defaults = {}
        self.atomtypes = {}
        self.atomstate = {}
        self.atom_name2atomnr = {}
        self.atomnr2atom_name = collections.defaultdict(list)
        self.nonbond_params = {}
        self.bondtypes, self.angletypes, self.dihedraltypes = {}, {}, {}
        self.moleculetype = {}
        self.molecules = []
        self.molecules_data = collections.defaultdict(dict)
        self.system_name = None
        self._mol = None

    # ---- reading
    def read(self):
        if self.content is None:
            with open(self.file_name) as f:
                self.content = f.readlines()
        section = prev = None
        for raw in self.content:
            line = re.sub(r";.*$", "", raw.strip())
            if not line or line[0] in "#;":
                continue
            if line.startswith("["):
                prev, section = section, line.strip("[] \t")
                if prev == "dihedrals" and section == "dihedrals":
                    section = "improper_dihedrals"
                continue
            handler = getattr(self, "_sec_" + (section or ""), None)
            if handler is not None:
                handler(line.split())
        return self

    def _sec_defaults(self, c):
        self.defaults = {"func": int(c[0]), "combinationrule": int(c[1]), "gen-pairs": len(c) > 2 and c[2] == "yes",
                         "fudgeLJ": float(c[3]) if len(c) > 3 else 1.0, "fudgeQQ": float(c[4]) if len(c) > 4 else 1.0, "nbfunc": 1}

    def _sec_atomtypes(self, c):
        # name [bond_type] [at.num] mass charge ptype sigma epsilon : 6, 7 or 8 (opls) columns
        if len(c) == 6:
            name, nr, mass, q, ptype, sig, eps = c[0], c[0], c[1], c[2], c[3], c[4], c[5]
        elif len(c) == 7:
            name, nr, mass, q, ptype, sig, eps = c[0], c[0], c[2], c[3], c[4], c[5], c[6]
        elif len(c) == 8 and c[0].startswith("opls"):
            name, nr, mass, q, ptype, sig, eps = c[0], c[1], c[3], c[4], c[5], c[6], c[7]
        else:
            print("Skip atom type %s" % c[0])
            return
        self.atom_name2atomnr[name] = nr
        self.atomnr2atom_name[nr].append(name)
        self.atomtypes[name] = {"name": name, "mass": float(mass), "charge": float(q), "type": ptype,
                                "sigma": float(sig), "epsilon": float(eps)}
        if name in self.atomstate:
            self.atomtypes[name]["state"] = self.atomstate[name]

    def _sec_atomstate(self, c):
        self.atomstate[c[0]] = int(c[1])
        if c[0] in self.atomtypes:
            self.atomtypes[c[0]]["state"] = int(c[1])

    def _sec_nonbond_params(self, c):
        key = tuple(sorted(c[:2]))
        if key in self.nonbond_params:
            raise RuntimeError("%s already exists, wrong topology" % (key,))
        self.nonbond_params[key] = {"func": int(c[2]), "params": c[3:]}

    def _sec_bondtypes(self, c):
        v = {"func": int(c[2]), "params": c[3:]}
        self.bondtypes.setdefault(c[0], {})[c[1]] = v
        self.bondtypes.setdefault(c[1], {})[c[0]] = v

    def _sec_angletypes(self, c):
        v = {"func": int(c[3]), "params": c[4:]}
        self.angletypes.setdefault(c[0], {}).setdefault(c[1], {})[c[2]] = v
        self.angletypes.setdefault(c[2], {}).setdefault(c[1], {})[c[0]] = v

    def _sec_dihedraltypes(self, c):
        try:
            v = {"func": int(c[4]), "params": c[5:]}
        except ValueError:
            print("Skip %s" % c)
            return
        self.dihedraltypes.setdefault(c[0], {}).setdefault(c[1], {}).setdefault(c[2], {})[c[3]] = v
        self.dihedraltypes.setdefault(c[3], {}).setdefault(c[2], {}).setdefault(c[1], {})[c[0]] = v

    def _sec_moleculetype(self, c):
        self._mol = c[0]
        self.moleculetype[c[0]] = int(c[1])

    def _need_mol(self):
        if self._mol is None:
            raise RuntimeError("Wrong order, before bonds there should be a moleculetype section")
        return self.molecules_data[self._mol]

    def _sec_atoms(self, c):
        a = TopoAtom()
        a.atom_id, a.atom_type, a.chain_idx, a.chain_name, a.name, a.cgnr = int(c[0]), c[1], int(c[2]), c[3], c[4], int(c[5])
        a.molecule_name = self._mol
        if len(c) > 6:
            a.charge = float(c[6])
        if len(c) > 7:
            a.mass = float(c[7])
        self._need_mol().setdefault("atoms", {})[a.atom_id] = a

    def _tuple_section(self, name, arity, c, skip=None):
        self._need_mol().setdefault(name, {})[tuple(int(x) for x in c[:arity])] = c[(skip or arity):]

    def _sec_bonds(self, c): self._tuple_section("bonds", 2, c)
    def _sec_angles(self, c): self._tuple_section("angles", 3, c)
    def _sec_dihedrals(self, c): self._tuple_section("dihedrals", 4, c)
    def _sec_improper_dihedrals(self, c): self._tuple_section("improper_dihedrals", 4, c, skip=3)
    def _sec_pairs(self, c): self._tuple_section("pairs", 2, c)

    def _sec_system(self, c):
        self.system_name = c[0]

    def _sec_molecules(self, c):
        self.molecules.append((c[0], int(c[1])))      # order matters

    # ---- writing (final `_output_topol.top`: src/start_simulation.py:834-994 + the section writers files_io.py:823-954)
    def write_system(self, path, atoms, bonds, angles, dihedrals, type_names):
        """The whole system as ONE molecule `MOL` (nrexcl 3), with the force-field sections of the input topology -- atomtypes,
        bondtypes, angletypes, dihedraltypes, nonbond_params, atomstate -- so that the file can be fed back to the driver.
        Tuple rows are `ids [func parameters ; origin]` as in the _bonds/_angles/_dihedrals.dat files, sorted like the reference's
        `_write_default`."""
        d = dict({"nbfunc": 1, "combinationrule": 1, "gen-pairs": False, "fudgeLJ": 1.0, "fudgeQQ": 1.0}, **self.defaults)
        sections = [("defaults", ["%s %s %s %s %s" % (d["nbfunc"], d["combinationrule"], "yes" if d["gen-pairs"] else "no", d["fudgeLJ"], d["fudgeQQ"])]),
                    ("atomtypes", ["%s %s %s %s %s %s" % (a["name"], a["mass"], a["charge"], a["type"], a["sigma"], a["epsilon"]) for a in self.atomtypes.values()])]

        def walk(node, prefix, depth):
            if depth == 0:
                yield prefix, node
            else:
                for k, v in node.items():
                    yield from walk(v, prefix + [k], depth - 1)
        for title, store, depth in (("bondtypes", self.bondtypes, 2), ("angletypes", self.angletypes, 3), ("dihedraltypes", self.dihedraltypes, 4)):
            rows = ["%s %s %s" % (" ".join(names), p["func"], " ".join(str(x) for x in p["params"])) for names, p in walk(store, [], depth)]
            if rows:
                sections.append((title, rows))
        if self.nonbond_params:
            sections.append(("nonbond_params", ["%s %s %s %s" % (k[0], k[1], p["func"], " ".join(str(x) for x in p["params"])) for k, p in self.nonbond_params.items()]))
        if self.atomstate:
            sections.append(("atomstate", ["%s %s" % kv for kv in self.atomstate.items()]))
        sections.append(("moleculetype", ["MOL 3"]))
        rows = []
        for aid in sorted(atoms):
            a = atoms[aid]
            rows.append("%s %s %s %s %s %s %s %s" % (aid, type_names[a["type_id"]], a["chain_idx"], a["chain_name"], a["name"], aid,
                                                     a["charge"] if a.get("charge") is not None else "0.0", a["mass"] if a.get("mass") is not None else ""))
        sections.append(("atoms", rows))
        for title, trows in (("bonds", bonds), ("angles", angles), ("dihedrals", dihedrals)):
            sections.append((title, [" ".join(str(x) for x in r) for r in sorted((list(r) for r in trows), key=lambda r: [x for x in r if isinstance(x, int)])]))
        sections += [("pairs", []), ("system", [self.system_name or "system"]), ("molecules", ["MOL 1"])]
        # layout of GROMACSTopologyFile.write (files_io.py:576-604): an empty line, the section name, its rows, an empty line
        with open(path, "w") as f:
            for title, rows in sections:
                f.write("\n[ %s ]\n" % title)
                f.writelines(r + "\n" for r in rows)
                f.write("\n")
