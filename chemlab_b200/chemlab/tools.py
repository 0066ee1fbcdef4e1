"""Driver helpers: conversion stop criteria and timer aggregation (src/tools.py:28-180)."""
import re

from .. import espressopp

_RE_TS = re.compile(r"(?P<type>[A-Za-z0-9-]+)\(?(?P<state>\d?)\)?")


class _SumObs:
    def __init__(self, parts, total):
        self.parts, self.total = parts, total

    def compute(self):
        return sum(p.compute() for p in self.parts)


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
        parts = []
        for s in sym.split("+"):
            m = _RE_TS.match(s).groupdict()
            state = int(m["state"]) if m["state"] else None
            key = (ids[m["type"]], tot, state)
            if key not in cr_observs:
                cr_observs[key] = (espressopp.analysis.ChemicalConversion(system, key[0], tot) if state is None
                                   else espressopp.analysis.ChemicalConversionTypeState(system, key[0], state, tot))
            parts.append(cr_observs[key])
        out.append((parts[0] if len(parts) == 1 else _SumObs(parts, tot), float(max_n) / tot))
    return out


def get_integrator_timers(system, integrator):
    """per-interaction timer labels like the reference's f<i> buckets (src/tools.py:51-79)."""
    e = system._ctx.engine
    if e is None:
        return {}
    t, c = e.timers()
    return dict(t, **{"n_" + k: v for k, v in c.items()})
