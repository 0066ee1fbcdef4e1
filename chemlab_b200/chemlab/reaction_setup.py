"""Reaction assembly: one ChemicalReaction extension, one FixedPairList + bond interaction per reaction group, one
Reaction object per [reaction_*] section with its post-processes (src/chemlab/reaction_setup.py:29-552,
src/chemlab/reaction_post_process.py:38-426).  Reverse / exchange / restricted reactions and the dummy-particle
post-processes are outside the engine's scope and raise NotImplementedError."""
import collections
import re

from .. import espressopp
from .reaction_parser import REACTION_NORMAL

EXT_POSTPROCESS, EXT_INTEGRATOR = "PP", "Integrator"
Ext = collections.namedtuple("Ext", "ext pp_type ext_type")
FPLDef = collections.namedtuple("FPLDef", "fpl type_list")


class PostProcessSetup:
    def __init__(self, system, topol, topol_manager, args):
        self.system, self.tm, self.topol, self.args = system, topol_manager, topol, args
        self.name2type = topol.atomsym_atomtype
        self.dynamic_types = set()
        self.observed_bondtypes = set()
        self.cr_observs = {}
        self.fix_distances = []
        self.use_thermal_group = False

    def setup(self, cfg):
        kind = cfg["class"]
        handler = {"ChangeNeighboursProperty": self._change_neighbour, "ATRPActivator": self._atrp_activator}.get(kind)
        if handler is None:
            raise NotImplementedError("reaction extension %s is outside the scope of the B200 engine (SURVEY E21)" % kind)
        return handler(cfg["options"])

    _RE_NEW = re.compile(r"(?P<type_name>\w+)\(?(?P<options>[a-zA-Z0-9_=,]*)\)?")

    def _change_neighbour(self, cfg):
        """type_transfers = OLD:nb_level->NEW[(state=..)] (examples/atrp_lj/atrp.cfg:8-14; reaction_post_process.py:78-115)."""
        pp = espressopp.integrator.PostProcessChangeNeighboursProperty(self.tm)
        for item in cfg["type_transfers"].split(","):
            old, new = item.split("->")
            old_type, level = old.split(":")
            m = self._RE_NEW.match(new.strip())
            new_name, options = m.group("type_name"), m.group("options")
            props = self.topol.gt.atomtypes[new_name]
            if "state" not in props:
                raise RuntimeError("Please define initial atom state in [ atomstate ] section of your topology for atom type %s" % new_name)
            kw = {"type": self.name2type[new_name], "mass": props["mass"], "q": props["charge"], "state": props["state"]}
            if options:
                for opt in options.split(","):
                    k, v = opt.split("=")
                    kw[k] = int(v) if k in ("state", "incr_state") else float(v)
            t_old, t_new = self.name2type[old_type.strip()], kw["type"]
            self.dynamic_types.update((t_old, t_new))
            pp.add_change_property(t_old, espressopp.integrator.TopologyParticleProperties(**kw), int(level))
        return Ext(pp, cfg.get("invoke_on"), EXT_POSTPROCESS)

    _RE_CENTER = re.compile(r"(?P<name>\w+)\((?P<state>\d+),\s*(?P<flag>[AD]{1,2})\)")
    _RE_PRODUCT = re.compile(r"(?P<new_type>\w+)\((?P<delta>[0-9-]+)\)")

    def _atrp_activator(self, cfg):
        """examples/atrp_lj/atrp.cfg:16-26; reaction_post_process.py:380-426."""
        a = self.args
        act = espressopp.integrator.ATRPActivator(self.system, int(cfg["interval"]), int(cfg["num_particles"]), float(cfg["ratio_activator"]),
                                                  float(cfg["ratio_deactivator"]), float(cfg["delta_catalyst"]), float(cfg["k_activate"]),
                                                  float(cfg["k_deactivate"]))
        act.stats_filename = cfg.get("stats_file", "%s_%s_atrp_stats.dat" % (a.output_prefix, a.rng_seed))
        act.select_from_all = int(cfg.get("select_from_all", 1))
        for opt in cfg["options"].split(";"):
            lhs, rhs = opt.split("->")
            c, p = self._RE_CENTER.match(lhs.strip()).groupdict(), self._RE_PRODUCT.match(rhs.strip()).groupdict()
            if c["flag"] not in ("A", "DA"):
                raise RuntimeError('Flag %s not "A" or "DA"' % c["flag"])
            prop = self.topol.gt.atomtypes[p["new_type"]]
            act.add_reactive_center(type_id=self.name2type[c["name"]], state=int(c["state"]), is_activator=(c["flag"] == "DA"),
                                    new_property=espressopp.integrator.TopologyParticleProperties(
                                        type=self.name2type[p["new_type"]], mass=prop["mass"], q=prop["charge"]),
                                    delta_state=int(p["delta"]))
        return Ext(act, None, EXT_INTEGRATOR)


class SetupReactions:
    def __init__(self, system, vl, topol, topol_manager, config, args):
        self.system, self.vl, self.topol, self.tm, self.cfg, self.args = system, vl, topol, topol_manager, config, args
        self.name2type = topol.atomsym_atomtype
        self.dynamic_types = set()
        self.observed_bondtypes = set()
        self.separate_fpls = set()
        self.cr_observs = {}
        self.fix_distances = []
        self.exclusions_list = []
        self.pp = PostProcessSetup(system, topol, topol_manager, args)
        self.pp.dynamic_types = self.dynamic_types
        self.reaction_index = {}

    use_thermal_group = property(lambda s: s.pp.use_thermal_group)

    def _reaction(self, r, fpl):
        """integrator.Reaction from one [reaction_*] section (reaction_setup.py:71-165)."""
        if r["reaction_type"] != REACTION_NORMAL or "sigma" in r:
            raise NotImplementedError("reaction %r: reverse/exchange reactions and random cut-offs are outside "
                                      "the scope of the B200 engine (SURVEY E21)" % r["equation"])
        rl = r["reactant_list"]
        a, b = rl["type_1"], rl["type_2"]
        t1, t2 = self.name2type[a["name"]], self.name2type[b["name"]]
        r_class = espressopp.integrator.RestrictReaction if r.get("connectivity_map") else espressopp.integrator.Reaction   # :74-77
        reaction = r_class(
            type_1=t1, type_2=t2, delta_1=int(a["delta"]), delta_2=int(b["delta"]), min_state_1=int(a["min"]), max_state_1=int(a["max"]),
            min_state_2=int(b["min"]), max_state_2=int(b["max"]), rate=float(r["rate"]), fpl=fpl, cutoff=float(r["cutoff"]))
        self.dynamic_types.update((t1, t2))
        reaction.intramolecular = bool(r["intramolecular"])
        reaction.intraresidual = bool(r["intraresidual"])
        reaction.is_virtual = bool(r["virtual"])
        if "min_cutoff" in r:
            reaction.get_reaction_cutoff().min_cutoff = float(r["min_cutoff"])
        reaction.active = r.get("active", True)
        if r.get("connectivity_map"):                                   # :115-126
            print("Reading connectivity map %s, reaction will be restricted to form connections only from the map" % r["connectivity_map"])
            ex_list = set()
            with open(r["connectivity_map"]) as f:
                for l in f:
                    if l.strip():
                        b1, b2 = (int(x) for x in l.split())
                        ex_list.add(tuple(sorted((b1, b2))))
            for b1, b2 in sorted(ex_list):
                reaction.define_connection(b1, b2)
            self.exclusions_list.extend(sorted(ex_list))
            print("Restricted to %d connections" % len(ex_list))
        n1, n2 = self.name2type[a["new_type"]], self.name2type[b["new_type"]]
        for side, old, new, new_name in (("type_1", t1, n1, a["new_type"]), ("type_2", t2, n2, b["new_type"])):
            if old != new:       # PostProcessChangeProperty: reactant type/mass/charge rewrite (:137-163)
                pp = espressopp.integrator.PostProcessChangeProperty()
                prop = self.topol.gt.atomtypes[new_name]
                pp.add_change_property(old, espressopp.integrator.TopologyParticleProperties(type=new, mass=prop["mass"], q=prop["charge"]))
                reaction.add_postprocess(pp, side)
                self.dynamic_types.update((old, new))
        return reaction, [(t1, t2), (n1, n2)]

    def setup_reactions(self):
        g = self.cfg["general"]
        self.ar_interval = int(g["interval"])
        ar = espressopp.integrator.ChemicalReaction(self.system, self.vl, self.system.storage, self.tm, self.ar_interval)
        ar.nearest_mode = g["nearest"]
        if g["pair_distances_filename"]:
            ar.pair_distances_filename = g["pair_distances_filename"]
        if g["max_per_interval"] > 0:
            ar.max_per_interval = g["max_per_interval"]
        fpls, reactions, to_integrator = [], [], []
        if getattr(self.args, "t_hybrid_bond", 0) > 0:
            raise NotImplementedError("hybrid (lambda) bonds are outside the scope of the B200 engine (SURVEY E21)")
        for group_name, group in self.cfg["reactions"].items():
            fpl = espressopp.FixedPairList(self.system.storage)
            pot_class = getattr(espressopp.interaction, group["potential"])
            opts = {}
            for k, v in group["potential_options"].items():
                try:
                    opts[k] = float(v)
                except ValueError:
                    opts[k] = v
            inter = getattr(espressopp.interaction, "FixedPairList%s" % group["potential"])(self.system, fpl, pot_class(**opts))
            fpl.interaction = inter
            self.system.addInteraction(inter, "chem_fpl_%s" % group_name)     # reaction bonds come first (reaction_setup.py:467)
            per_reaction = collections.defaultdict(list)
            to_integrator = []      # re-initialised for every group, as in the reference (:471): the integrator extensions (ATRPActivator)
            #                         of the LAST group are the ones the driver adds
            for ext_name, ext_cfg in group["extensions"].items():
                x = self.pp.setup(ext_cfg)
                if x.ext_type == EXT_INTEGRATOR:
                    to_integrator.append(x.ext)
                else:
                    per_reaction[ext_name].append(x)
            type_list = []
            for r in group["reaction_list"]:
                r["connectivity_map"] = group["connectivity_map"]
                reaction, types = self._reaction(r, fpl)
                type_list.extend(types)
                for ext_name, exts in per_reaction.items():
                    if ext_name in r["exclude_extensions"]:
                        continue
                    for x in exts:
                        reaction.add_postprocess(x.ext, x.pp_type) if x.pp_type else reaction.add_postprocess(x.ext)
                ar.add_reaction(reaction)
                self.reaction_index[len(reactions)] = r["equation"]
                reactions.append(reaction)
            fpls.append(FPLDef(fpl, set(type_list)))
        return ar, fpls, reactions, to_integrator

    def rebuild_fixed_pair_lists(self):
        pass
