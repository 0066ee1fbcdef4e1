"""Iteration order of the plain dicts of the Python-2 reference, where that order is visible in chemlab's results: the type ids of
the atom types (src/chemlab/gromacs_topology.py:260) and the order of the reaction groups (src/chemlab/reaction_parser.py:239,
src/chemlab/reaction_setup.py:434).  Pinned by the shipped run log examples/atrp_lj/single:199-205 (tests/test_runlog_cpu.py)."""


def _py2_str_hash(s):
    """hash() of a str in CPython 2.7 on a 64-bit build without -R (Objects/stringobject.c:string_hash), as unsigned 64 bit."""
    if not s:
        return 0
    m = (1 << 64) - 1
    x = (ord(s[0]) << 7) & m
    for ch in s:
        x = ((1000003 * x) & m) ^ ord(ch)
    x ^= len(s)
    return m - 1 if x == m else x


def py2_dict_order(keys):
    """Iteration order of a CPython-2.7 dict into which the str `keys` were inserted in this order (no deletions):
    open addressing with the perturbed probe sequence of Objects/dictobject.c, 8 slots at first, resize to the first power of two
    above 4 * used once two thirds are filled; iteration runs over the slots.  chemlab's type ids depend on it (see _prepare)."""
    import os
    if os.environ.get("CHEMLAB_PY2_ORDER", "1") == "0":      # insertion order: what the reference's code does when it runs under Python 3
        seen, out = set(), []                              # (tests/test_reference_driver_cpu.py compares with exactly that)
        for k in keys:
            if k not in seen:
                seen.add(k); out.append(k)
        return out

    def insert(table, k, h):
        mask = len(table) - 1
        i, perturb = h & mask, h
        while table[i & mask] is not None:
            if table[i & mask][0] == k:
                return False
            i = (i << 2) + i + perturb + 1
            perturb >>= 5
        table[i & mask] = (k, h)
        return True
    table, fill = [None] * 8, 0
    for k in keys:
        if insert(table, k, _py2_str_hash(k)):
            fill += 1
            if fill * 3 >= len(table) * 2:
                size = 8
                while size <= (2 if fill > 50000 else 4) * fill:
                    size <<= 1
                old, table = [e for e in table if e is not None], [None] * size
                for kk, hh in old:
                    insert(table, kk, hh)
    return [e[0] for e in table if e is not None]
