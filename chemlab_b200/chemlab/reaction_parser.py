"""Reaction `.cfg` (INI) parser: sections [general], [group_X], [ext_Y], [reaction_Z] and the equation grammars
`A(min,max) + B(min,max) -> C(delta):D(delta)` (src/chemlab/reaction_parser.py:36-266)."""
import configparser
import re

REACTION_NORMAL, REACTION_DISSOCIATION, REACTION_EXCHANGE = "normal", "diss", "exchange"

_REACTANT = r"(?P<name>\w+)\((?P<min>\d+),\s*(?P<max>\d+)\)"
_PRODUCT = r"(?P<name>\w+)\((?P<delta>[0-9-]+)\)"
_RE_REACTANT, _RE_PRODUCT = re.compile(_REACTANT), re.compile(_PRODUCT)


def _reactant(s):
    m = _RE_REACTANT.fullmatch(s.strip())
    if not m:
        raise ValueError("bad reactant %r" % s)
    return m.groupdict()


def _product(s):
    m = _RE_PRODUCT.fullmatch(s.strip())
    if not m:
        raise ValueError("bad product %r" % s)
    return m.groupdict()


def parse_equation(eq):
    """`A(1,2) + B(0,3) -> C(1):D(-1)`  (bond formation)."""
    lhs, rhs = eq.split("->")
    a, b = (_reactant(x) for x in lhs.split("+"))
    pa, pb = (_product(x) for x in rhs.split(":"))
    out = {"type_1": a, "type_2": b}
    for side, prod in (("type_1", pa), ("type_2", pb)):
        out[side]["delta"] = prod["delta"]
        out[side]["new_type"] = prod["name"]
    return out, REACTION_NORMAL


def parse_reverse_equation(eq):
    """`A(1,2):B(0,3) -> C(1) + D(-1)`  (bond dissociation)."""
    lhs, rhs = eq.split("->")
    a, b = (_reactant(x) for x in lhs.split(":"))
    pa, pb = (_product(x) for x in rhs.split("+"))
    out = {"type_1": a, "type_2": b}
    for side, prod in (("type_1", pa), ("type_2", pb)):
        out[side]["delta"] = prod["delta"]
        out[side]["new_type"] = prod["name"]
    return out, REACTION_DISSOCIATION


def parse_exchange_equation(eq):
    """`A(0,1):B(0,1) + C(0,1) -> D(1):E(1) + F(1)`  (bond exchange; src/tests/test_reaction_parser.py:29-51)."""
    lhs, rhs = eq.split("->")
    bonded, free = lhs.split("+")
    a, b = (_reactant(x) for x in bonded.split(":"))
    c = _reactant(free)
    new_bond, released = rhs.split("+")
    pa, pc = (_product(x) for x in new_bond.split(":"))
    pb = _product(released)
    out = {"type_1": a, "type_2": b, "type_3": c}
    for side, prod in (("type_1", pa), ("type_2", pb), ("type_3", pc)):
        out[side]["delta"] = prod["delta"]
        out[side]["new_type"] = prod["name"]
    return out, REACTION_EXCHANGE


def _literal(s, default):
    import ast
    return ast.literal_eval(s) if s is not None else default


def process_reaction(items):
    r = dict(items)
    data = {"rate": float(r["rate"]), "intramolecular": _literal(r.get("intramolecular"), False),
            "intraresidual": _literal(r.get("intraresidual"), False), "virtual": _literal(r.get("virtual"), False),
            "exclude_extensions": [], "equation": r["reaction"], "active": _literal(r.get("active"), True)}
    if "exclude_extensions" in r:
        data["exclude_extensions"] = {s.strip() for s in r["exclude_extensions"].split(",")}
    kind = None
    for parser in (parse_equation, parse_reverse_equation, parse_exchange_equation):
        try:
            data["reactant_list"], kind = parser(r["reaction"])
            break
        except (ValueError, AttributeError):
            continue
    if kind is None:
        raise RuntimeError("Could not parse reaction equation: %s" % r["reaction"])
    data["reaction_type"] = kind
    if "min_cutoff" in r:
        data["min_cutoff"] = float(r["min_cutoff"])
    if "sigma" in r and "eq_distance" in r:
        data["sigma"], data["eq_distance"] = float(r["sigma"]), float(r["eq_distance"])
    elif "cutoff" in r:
        data["cutoff"] = float(r["cutoff"])
    else:
        raise RuntimeError("Please define cutoff of the reaction: %s" % r["reaction"])
    if kind == REACTION_DISSOCIATION:
        if "diss_rate" in r:
            data["diss_rate"] = float(r["diss_rate"])
        data["alpha"] = float(r["alpha"])
    return r["group"], data


def process_general(items):
    g = dict(items)
    # NB: the reference evaluates bool() on the raw string, so `nearest=0` means True; preserved on purpose
    return {"interval": int(g["interval"]), "nearest": bool(g.get("nearest", False)),
            "pair_distances_filename": g.get("pair_distances_filename"), "max_per_interval": int(g.get("max_per_interval", -1))}


def process_group(items):
    g = dict(items)
    out = {"reaction_list": [], "connectivity_map": g.get("connectivity_map"), "extensions": {}}
    if "extensions" in g:
        out["extensions"] = {s.strip(): None for s in g["extensions"].split(",")}
    if "potential" in g:
        out["potential"] = g["potential"]
        out["potential_options"] = dict(s.split("=") for s in g["potential_options"].split(","))
    return out


def parse_config(path):
    # strict=False: Python 2's ConfigParser (the reference) let a repeated option overwrite the earlier one
    cp = configparser.ConfigParser(interpolation=None, strict=False)
    cp.optionxform = str.lower
    if not cp.read(path):
        raise RuntimeError("cannot read reaction config %s" % path)
    config, extensions = {"general": None, "reactions": {}}, {}
    for s in cp.sections():
        items = list(cp.items(s))
        if s == "general":
            config["general"] = process_general(items)
        elif s.startswith("ext_"):
            name = s[4:].strip()
            if name in extensions:
                raise RuntimeError("Name of extension already exists")
            d = dict(items)
            extensions[name] = {"class": d.pop("ext_type"), "options": d}
        elif s.startswith("group_"):
            name = s[6:].strip()
            if name not in config["reactions"]:
                grp = process_group(items)
                for ext in grp["extensions"]:
                    grp["extensions"][ext] = extensions[ext]     # extensions must precede the groups that name them
                config["reactions"][name] = grp
        elif s.startswith("reaction_"):
            group, data = process_reaction(items)
            if group not in config["reactions"]:
                raise RuntimeError("Wrong order, first reaction groups and then referring reactions")
            config["reactions"][group]["reaction_list"].append(data)
    # the reference keeps the groups in a plain Python-2 dict (reaction_parser.py:239) and SetupReactions walks `.items()`
    # (reaction_setup.py:434): the order of the reaction lists (chem_fpl index, count_<k> column, reaction index) is CPython 2.7's
    # hash order of the group names -- e.g. reaction_2 before reaction_1 for examples/rim135/reaction.cfg
    from .py2compat import py2_dict_order
    config["reactions"] = {k: config["reactions"][k] for k in py2_dict_order(list(config["reactions"]))}
    return config
