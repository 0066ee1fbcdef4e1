"""ctypes front-end of the fp64 CPU oracle (oracle/chemlab_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package `chemlab_b200` never imports this.
PARITY UNPINNED (see the header of chemlab_oracle.c and REFERENCE_UNVERIFIED.md).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "chemlab_oracle.c")


def build(force=False):
    """Compile the C restatement with gcc (-O2 -fopenmp). Idempotent."""
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    # -msse4.1: nearbyint() of the minimum image becomes one ROUNDSD instead of a libm call (same rounding, no FMA contraction)
    cmd = ["gcc", "-O2", "-msse4.1", "-fopenmp", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_uint64]
        for name in ("orc_list_size", "orc_pairs_brute", "orc_get_pairs", "orc_get_pairs_raw", "orc_list_get", "orc_step",
                     "orc_nrebuild", "orc_npairs", "orc_count_interacting", "orc_reaction_counter",
                     "orc_last_events", "orc_get_candidates", "orc_count_type", "orc_get_exclusions"):
            getattr(L, name).restype = C.c_int64
        L.orc_energy.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Oracle:
    """Dense-index (0..n-1) fp64 simulator mirroring the engine's C-ABI one call at a time."""

    def __init__(self, n, box, rc, skin, seed=0):
        self.L = lib()
        self.n = int(n)
        b = _d(box)
        self.h = C.c_void_p(self.L.orc_create(self.n, _p(b), C.c_double(rc), C.c_double(skin), C.c_uint64(seed)))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # -- particles
    def set_particles(self, pos, vel, mass, q, type, state=None, resid=None):
        n = self.n
        pos = _d(pos).reshape(n, 3)
        vel = _d(vel).reshape(n, 3) if vel is not None else None
        mass = _d(mass)
        q = _d(q) if q is not None else None
        type = _i(type)
        state = _i(state) if state is not None else None
        resid = _i(resid) if resid is not None else None
        self.L.orc_set_particles(self.h, _p(pos), _p(vel), _p(mass), _p(q), _p(type, C.c_int),
                                 _p(state, C.c_int), _p(resid, C.c_int))

    def set_positions(self, pos):
        pos = _d(pos)
        self.L.orc_set_positions(self.h, _p(pos))

    def set_velocities(self, vel):
        vel = _d(vel)
        self.L.orc_set_velocities(self.h, _p(vel))

    def get(self):
        n = self.n
        x = np.zeros((n, 3)); v = np.zeros((n, 3)); f = np.zeros((n, 3))
        ty = np.zeros(n, np.int32); st = np.zeros(n, np.int32); m = np.zeros(n)
        im = np.zeros((n, 3), np.int32)
        self.L.orc_get(self.h, _p(x), _p(v), _p(f), _p(ty, C.c_int), _p(st, C.c_int), _p(m), _p(im, C.c_int))
        return dict(pos=x, vel=v, force=f, type=ty, state=st, mass=m, image=im)

    def modify(self, i, field, value):
        v = _d(np.atleast_1d(value))
        self.L.orc_modify(self.h, int(i), int(field), _p(v))

    def set_option(self, name, value):
        self.L.orc_set_option(self.h, name.encode(), C.c_double(value))

    def set_threads(self, nt):
        self.L.orc_set_threads(self.h, int(nt))

    def max_threads(self):
        return int(self.L.orc_max_threads())

    # -- exclusions / tables / potentials
    def set_exclusions(self, pairs):
        pairs = _i(pairs).reshape(-1, 2)
        self.L.orc_set_exclusions(self.h, C.c_int64(len(pairs)), _p(pairs, C.c_int))

    def get_exclusions(self):
        n = self.L.orc_get_exclusions(self.h, C.c_int64(0), None)
        out = np.zeros((n, 2), np.int32)
        self.L.orc_get_exclusions(self.h, C.c_int64(n), _p(out, C.c_int))
        return out

    def excl_observe(self, lst):
        self.L.orc_excl_observe(self.h, int(lst))

    def add_table(self, x, e, f, interp=1):
        x, e, f = _d(x), _d(e), _d(f)
        return int(self.L.orc_add_table(self.h, len(x), _p(x), _p(e), _p(f), int(interp)))

    def table_eval(self, tab, x):
        e = C.c_double(); f = C.c_double()
        bad = self.L.orc_table_eval(self.h, int(tab), C.c_double(x), C.byref(e), C.byref(f))
        return e.value, f.value, bad

    def add_nonbonded(self, kind):
        return int(self.L.orc_add_interaction(self.h, int(kind)))

    def nb_set_tab(self, inter, t1, t2, tab, rc):
        self.L.orc_nb_set_tab(self.h, inter, t1, t2, tab, C.c_double(rc))

    def nb_set_lj(self, inter, t1, t2, eps, sig, rc, shift_auto=1):
        self.L.orc_nb_set_lj(self.h, inter, t1, t2, C.c_double(eps), C.c_double(sig), C.c_double(rc), int(shift_auto))

    def nb_set_mixed(self, inter, t1, t2, tab1, tab2, mix, conv_type, conv_total, rc):
        self.L.orc_nb_set_mixed(self.h, inter, t1, t2, tab1, tab2, C.c_double(mix), int(conv_type),
                                C.c_double(conv_total), C.c_double(rc))

    # -- lists / bonded
    def add_list(self, arity):
        return int(self.L.orc_add_list(self.h, int(arity)))

    def list_add(self, lst, ids):
        ids = _i(ids)
        ar = ids.shape[-1] if ids.ndim > 1 else None
        n = len(ids) if ar else 0
        self.L.orc_list_add(self.h, int(lst), C.c_int64(n), _p(ids, C.c_int))

    def list_size(self, lst):
        return int(self.L.orc_list_size(self.h, int(lst)))

    def list_get(self, lst, arity):
        n = self.list_size(lst)
        out = np.zeros((n, arity), np.int32)
        self.L.orc_list_get(self.h, int(lst), C.c_int64(n), _p(out, C.c_int))
        return out

    def add_bonded(self, lst, typed=0):
        return int(self.L.orc_add_bonded(self.h, int(lst), int(typed)))

    def bonded_set_potential(self, inter, types, kind, params=(), table=-1):
        t = list(types) + [-1] * (4 - len(types))
        p = _d(list(params) if len(params) else [0.0])
        self.L.orc_bonded_set_potential(self.h, inter, t[0], t[1], t[2], t[3], int(kind), _p(p), len(params), int(table))

    # -- lists / forces
    def rebuild(self):
        self.L.orc_rebuild(self.h)

    def pairs(self):
        n = self.L.orc_get_pairs(self.h, C.c_int64(0), None)
        out = np.zeros((n, 2), np.int32)
        self.L.orc_get_pairs(self.h, C.c_int64(n), _p(out, C.c_int))
        return out

    def pairs_raw(self):
        """Pair set in list order (rows (min, max), unsorted)."""
        n = self.L.orc_get_pairs_raw(self.h, C.c_int64(0), None)
        out = np.zeros((n, 2), np.int32)
        self.L.orc_get_pairs_raw(self.h, C.c_int64(n), _p(out, C.c_int))
        return out

    def pairs_brute(self):
        n = self.L.orc_pairs_brute(self.h, C.c_int64(0), None)
        out = np.zeros((n, 2), np.int32)
        self.L.orc_pairs_brute(self.h, C.c_int64(n), _p(out, C.c_int))
        return out

    def npairs(self):
        return int(self.L.orc_npairs(self.h))

    def count_interacting(self):
        return int(self.L.orc_count_interacting(self.h))

    def compute_forces(self):
        self.L.orc_compute_forces(self.h)

    def energy(self, inter):
        return float(self.L.orc_energy(self.h, int(inter)))

    def kinetics(self):
        out = np.zeros(3)
        self.L.orc_kinetics(self.h, _p(out))
        return out

    def range_error(self):
        return int(self.L.orc_range_error(self.h))

    # -- integrator
    def set_dt(self, dt):
        self.L.orc_set_dt(self.h, C.c_double(dt))

    def set_langevin(self, on, kT, gamma, types=()):
        t = _i(list(types) if len(types) else [0])
        self.L.orc_set_langevin(self.h, int(on), C.c_double(kT), C.c_double(gamma), len(types), _p(t, C.c_int))

    def run(self, n):
        self.L.orc_run(self.h, C.c_int64(n))

    def run_continue(self, n):
        self.L.orc_run_continue(self.h, C.c_int64(n))

    def step(self):
        return int(self.L.orc_step(self.h))

    def nrebuild(self):
        return int(self.L.orc_nrebuild(self.h))

    # -- reactions
    def reaction_general(self, on, interval, nearest, max_per_interval=0):
        self.L.orc_reaction_general(self.h, int(on), int(interval), int(nearest), int(max_per_interval))

    def add_reaction(self, type_1, type_2, delta_1, delta_2, min1, max1, min2, max2, rate, cutoff,
                     lst, min_cutoff=0.0, intramolecular=1, intraresidual=1, is_virtual=0, active=1):
        return int(self.L.orc_add_reaction(self.h, type_1, type_2, delta_1, delta_2, min1, max1, min2, max2,
                                           C.c_double(rate), C.c_double(cutoff), C.c_double(min_cutoff), int(lst),
                                           int(intramolecular), int(intraresidual), int(is_virtual), int(active)))

    def reaction_define_connections(self, r, pairs):
        p = _i(np.asarray(pairs).reshape(-1, 2))
        self.L.orc_reaction_define_connections(self.h, int(r), C.c_int64(len(p)), _p(p, C.c_int))

    def reaction_set_rate(self, r, rate):
        self.L.orc_reaction_set_rate(self.h, int(r), C.c_double(rate))

    def reaction_set_active(self, r, a):
        self.L.orc_reaction_set_active(self.h, int(r), int(a))

    def reaction_add_change(self, reaction, side, nb_level, old_type, new_type, new_mass=-1.0,
                            new_q=float("nan"), state_mode=0, state_value=0):
        self.L.orc_reaction_add_change(self.h, int(reaction), int(side), int(nb_level), int(old_type), int(new_type),
                                       C.c_double(new_mass), C.c_double(new_q), int(state_mode), int(state_value))

    def tm_observe(self, lst):
        self.L.orc_tm_observe(self.h, int(lst))

    def tm_register(self, lst, types):
        t = list(types) + [-1] * (4 - len(types))
        self.L.orc_tm_register(self.h, int(lst), t[0], t[1], t[2], t[3])

    def tm_initialize(self):
        self.L.orc_tm_initialize(self.h)

    def react(self):
        self.L.orc_react(self.h)
        return int(self.L.orc_last_events(self.h))

    def reaction_counter(self, r):
        return int(self.L.orc_reaction_counter(self.h, int(r)))

    def candidates(self):
        n = self.L.orc_get_candidates(self.h, C.c_int64(0), None, None)
        rows = np.zeros((n, 4), np.int32); d2 = np.zeros(n)
        self.L.orc_get_candidates(self.h, C.c_int64(n), _p(rows, C.c_int), _p(d2))
        return rows, d2

    def count_type(self, t, state=-1):
        return int(self.L.orc_count_type(self.h, int(t), int(state)))

    def set_cap_force(self, cap):
        self.L.orc_set_cap_force(self.h, C.c_double(cap))

    # -- ATRPActivator
    def atrp_configure(self, num_particles, ratio_activator, ratio_deactivator, delta_catalyst, k_activate, k_deactivate):
        self.L.orc_atrp_configure(self.h, int(num_particles), C.c_double(ratio_activator), C.c_double(ratio_deactivator),
                                  C.c_double(delta_catalyst), C.c_double(k_activate), C.c_double(k_deactivate))

    def atrp_add_center(self, type_id, state, needs_deactivator, new_type=-1, new_mass=-1.0, new_q=float("nan"), delta_state=0):
        self.L.orc_atrp_add_center(self.h, int(type_id), int(state), int(bool(needs_deactivator)), int(new_type), C.c_double(new_mass),
                                   C.c_double(new_q), int(delta_state))

    def atrp_now(self):
        cnt = (C.c_int64 * 2)(); rat = (C.c_double * 2)()
        self.L.orc_atrp_now(self.h, cnt, rat)
        return (int(cnt[0]), int(cnt[1])), (float(rat[0]), float(rat[1]))


# Engine-compatible method names, so that one set-up routine can drive either side
Oracle.nb_set_tabulated = Oracle.nb_set_tab
Oracle.exclusions_observe = Oracle.excl_observe
Oracle.topology_observe = Oracle.tm_observe
Oracle.topology_register = Oracle.tm_register
Oracle.topology_initialize = Oracle.tm_initialize
Oracle.react_now = Oracle.react


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox(c, k, o)
    return [int(x) for x in o]
