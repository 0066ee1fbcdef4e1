"""Command-line / @params interface of start_simulation.py (option names, defaults and the `key=value` arg-file
convention follow src/app_args.py:29-211 so that shipped `params` files work unchanged)."""
import argparse
import ast
import random

_lit = ast.literal_eval

# (group, [(flags, kwargs), ...]) -- one table instead of sixty add_argument calls
OPTIONS = [
    ("General options", [
        (("--conf",), dict(required=True)), (("--top", "--topology"), dict(required=True, dest="top")),
        (("--node_grid",), {}), (("--skin",), dict(default=0.16)), (("--output_prefix",), dict(default="sim")),
        (("--output_file",), dict(default="trjout.h5")), (("--trj_collect",), dict(default=1000, type=int)),
        (("--energy_collect",), dict(default=1000, type=int)), (("--topol_collect",), dict(default=1000, type=int)),
        (("--reactions",), dict(default=None)), (("--debug",), dict(default=None)),
        (("--check_topology",), dict(default=False, type=_lit)), (("--start_ar",), dict(default=0, type=int)),
        (("--stop_ar",), dict(default=-1, type=int)), (("--table_groups",), dict(default=None)),
        (("--max_force",), dict(default=-1, type=float)), (("--rate_arrhenius",), dict(default=False, type=_lit)),
        (("--exclusion_list",), dict(default=None)), (("--benchmark_data",), dict(default=None)),
        (("--system_monitor_filter",), dict(default=None)), (("--do_not_exclude_bonds",), dict(default=False, type=_lit)),
    ]),
    ("Simulation parameters", [
        (("--kb",), dict(type=float, default=0.0083144621)), (("--mass_factor",), dict(type=float, default=1.6605402)),
        (("--run",), dict(type=int, default=10000)), (("--int_step",), dict(type=int, default=1000)),
        (("--rng_seed",), dict(type=int, default=None)), (("--thermal_groups",), dict(default=None)),
        (("--gen_velocity",), dict(default=False, type=_lit)),
        (("--thermostat",), dict(default="lv", choices=("lv", "vr", "iso", "br", "no"))),
        (("--barostat",), dict(default="lv", choices=("lv", "br"))), (("--barostat_tau",), dict(default=5.0, type=float)),
        (("--barostat_mass",), dict(default=50.0, type=float)), (("--barostat_gammaP",), dict(default=1.0, type=float)),
        (("--thermostat_gamma",), dict(default=5.0, type=float)), (("--temperature",), dict(default=458.0, type=float)),
        (("--pressure",), dict(default=None, type=float)), (("--dt",), dict(default=0.001, type=float)),
        (("--lj_cutoff",), dict(default=1.2, type=float)), (("--cg_cutoff",), dict(default=1.4, type=float)),
        (("--coulomb_epsilon1",), dict(default=1.0, type=float)), (("--coulomb_epsilon2",), dict(default=80.0, type=float)),
        (("--coulomb_kappa",), dict(default=0.0, type=float)), (("--coulomb_cutoff",), dict(default=0.9, type=float)),
    ]),
    ("H5MD storage", [
        ((("--store_" + n),), dict(default=d, type=_lit)) for n, d in
        (("species", True), ("state", True), ("position", True), ("lambda", False), ("force", False), ("velocity", False),
         ("charge", False), ("mass", True), ("res_id", True), ("pressure", False), ("single_precision", True), ("angdih", False))
    ] + [(("--save_before_reaction",), dict(default=False, type=_lit)), (("--trj_flush",), dict(default=None, type=int)),
         (("--gro_trj_collect",), dict(default=None, type=int))]),
    ("Maximum conversion", [
        (("--maximum_conversion",), dict(default=None)), (("--eq_steps",), dict(default=0, type=int)),
        (("--keep_simulation",), dict(default=False)),
    ]),
    ("Counters", [
        (("--count_types",), dict(default=None)), (("--count_tuples",), dict(default=False, type=_lit)),
        (("--count_types_state",), dict(default=None)), (("--count_fix_distances",), dict(default=False, type=_lit)),
    ]),
    ("Hybrid bonded terms", [
        (("--t_hybrid_bond",), dict(default=0, type=int)), (("--t_hybrid_angle",), dict(default=0, type=int)),
        (("--t_hybrid_dihedral",), dict(default=0, type=int)),
    ]),
]


class ArgFileParser(argparse.ArgumentParser):
    """`@params` files hold one `key=value` per line; a bare key gets `--` prepended, `#`/`;` start a comment."""
    def convert_arg_line_to_args(self, line):
        out = []
        for tok in line.split():
            if tok[0] in "#;":
                break
            out.append(tok if tok.startswith("--") else "--" + tok)
        return out

    @staticmethod
    def save_to_file(path, namespace):
        with open(path, "w") as f:
            for k in sorted(vars(namespace)):
                v = getattr(namespace, k)
                if v is not None:
                    f.write("%s=%s\n" % (k, v))


def _args():
    p = ArgFileParser(description="Runs classical (reactive) MD simulation on the B200 engine", fromfile_prefix_chars="@")
    for title, opts in OPTIONS:
        g = p.add_argument_group(title)
        for flags, kw in opts:
            g.add_argument(*flags, **kw)
    return p


def parse(argv=None):
    a = _args().parse_args(argv)
    if a.rng_seed is None or a.rng_seed == -1:
        a.rng_seed = random.randint(1000, 10000)
    return a


class RegexpFilter:
    """logging filter of `--debug name:regex` (src/app_args.py:60-68): a record passes when its message or the name of the
    function that logged it matches the expression."""
    def __init__(self, regexp):
        import re
        self._rx = re.compile(regexp)

    def filter(self, record):
        return bool(self._rx.match(str(record.msg)) or self._rx.match(record.funcName or ""))
