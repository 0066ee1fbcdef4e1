"""Driver helpers: conversion stop criteria and timer aggregation (src/tools.py:28-180)."""
import re

from .. import espressopp

_RE_TS = re.compile(r"(?P<type>[A-Za-z0-9-]+)\(?(?P<state>\d?)\)?")


def get_maximum_conversion(args, system, chem_fpls, gt, cr_observs=None):
    """--maximum_conversion 'SYM(state):max:total[,…]' ; 'A-B:max:total' counts bonds of a reaction list; 'A+B:…' sums types
    (src/tools.py:102-180).  Returns [(observable, stop_value)]."""
    cr_observs = {} if cr_observs is None else cr_observs
    out = []
    ids = gt.used_atomsym_atomtype
    for item in args.maximum_conversion.split(","):
        sym, max_n, tot = item.split(":")
        max_n, tot = int(max_n), int(tot)
        if "-" in sym:
            a, b = _RE_TS.match(sym).group("type").split("-")
            for f in chem_fpls:
                if (ids[a], ids[b]) in f.type_list or (ids[b], ids[a]) in f.type_list:
                    out.append((espressopp.analysis.NFixedPairListEntries(system, f.fpl), max_n))
                    break
            continue
        if "+" in sym:                                        # :141-158: one observable counting every summand, column cr_<A>_<B>_...
            obs = espressopp.analysis.ChemicalConversionTypeState(system, total_count=tot)
            names = []
            for s in sym.split("+"):
                m = _RE_TS.match(s).groupdict()
                obs.count_type(ids[m["type"]], int(m["state"]) if m["state"] else None)
                names.append(m["type"])
            cr_observs[(tuple(names), tot, None)] = obs
            out.append((obs, float(max_n) / tot))
            continue
        m = _RE_TS.match(sym).groupdict()                      # :159-178
        state = int(m["state"]) if m["state"] else None
        key = (ids[m["type"]], tot, state)
        if key not in cr_observs:
            cr_observs[key] = (espressopp.analysis.ChemicalConversion(system, key[0], tot) if state is None
                               else espressopp.analysis.ChemicalConversionTypeState(system, key[0], state, tot))
        out.append((cr_observs[key], float(max_n) / tot))
    return out


def get_integrator_timers(system, integrator):
    """per-interaction timer labels like the reference's f<i> buckets (src/tools.py:51-79)."""
    e = system._ctx.engine
    if e is None:
        return {}
    t, c = e.timers()
    return dict(t, **{"n_" + k: v for k, v in c.items()})
