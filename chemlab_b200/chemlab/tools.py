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


def get_integrator_timers(alltimers, system):
    """integrator.getTimers() = one list of (name, seconds) pairs per rank -> {label: seconds averaged over the ranks}; 'timeRun' is
    skipped and 'f<i>' becomes the label of interaction i (src/tools.py:51-79)."""
    if not alltimers or not alltimers[0]:
        return {}
    nprocs = len(alltimers)
    timers = {k: 0.0 for k, _ in alltimers[0]}
    for ntimer in alltimers:
        for k, v in ntimer:
            if k != "timeRun":
                timers[k] += float(v)
    for k in timers:
        timers[k] /= nprocs
    for k, v in sorted(timers.items()):
        if k.startswith("f") and k[1:].isdigit():
            timers[system.getNameOfInteraction(int(k[1:]))] = v
            del timers[k]
    return timers


def average_timers(timer_list):
    """list per rank of (name, value) pairs -> {name: mean over the ranks} (src/tools.py:82-99)"""
    import collections
    acc = collections.defaultdict(list)
    for cpu_list in timer_list:
        for k, v in cpu_list:
            acc[k].append(v)
    return {k: sum(v) / float(len(v)) for k, v in acc.items() if v}
