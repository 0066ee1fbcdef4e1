"""Engine -- thin object wrapper over the C-ABI (include/chemlab_b200.h).

Every method is one C call; numpy arrays cross the boundary as plain pointers.  The espressopp-style
surface (chemlab_b200.espressopp) is built on this class.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ReactionSpec, c_f64p, c_i32p, c_i64p


class EngineError(RuntimeError):
    """Raised for every non-zero C-ABI status (reference behaviour: RuntimeError / MPI abort)."""


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


# potential kinds (CLB_POT_*), non-bonded kinds (CLB_NB_*)
POT = dict(Harmonic=1, Tabulated=2, AngularHarmonic=3, TabulatedAngular=4, TabulatedDihedral=5, Cosine=6, FENE=7,
           DihedralHarmonic=8, FENELennardJones=9, LennardJones=10)
NB = dict(Tabulated=1, LennardJones=2, MixedTabulated=3)


def slab_planes(ncz, nranks):
    """[(cz0, nczl)] per rank: the cell planes along z are dealt out contiguously, the first ncz % nranks ranks
    get one more (same rule as slab_planes() in csrc/engine_comm.inl)."""
    base, rem = divmod(int(ncz), int(nranks))
    out = []
    for r in range(nranks):
        out.append((r * base + min(r, rem), base + (1 if r < rem else 0)))
    return out


def slab_owner(z, box_z, rc_plus_skin, nranks):
    """Rank that owns coordinate(s) z: plane = floor(frac(z / Lz) * ncz) with ncz = floor(Lz / (rc + skin))."""
    ncz = int(np.floor(box_z / rc_plus_skin))
    fr = np.asarray(z, float) / box_z
    cz = np.minimum((np.floor((fr - np.floor(fr)) * ncz)).astype(np.int64), ncz - 1)
    bounds = np.array([c0 for c0, _ in slab_planes(ncz, nranks)] + [ncz])
    return np.searchsorted(bounds, cz, side="right") - 1


def broadcast_nccl_id(group=None):
    """128-byte ncclUniqueId created on rank 0 and broadcast over the torch.distributed group (nccl or gloo backend)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(Engine.nccl_unique_id()), dtype=torch.uint8).clone()
    buf = buf.to(dev)
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


class Engine:
    def __init__(self, box, rc_max, skin, seed=0, device=0):
        self.L = _lib.load()
        self.h = C.c_void_p()
        b = _f64(box)
        rc = self.L.clb_create(C.byref(self.h), int(device), _p(b, c_f64p), float(rc_max), float(skin), int(seed))
        if rc != 0:
            raise EngineError(self.L.clb_last_error(None).decode())
        self.box = np.array(b)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.clb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise EngineError("%s (status %d)" % (self.L.clb_last_error(self.h).decode(), rc))

    # ---- options
    def set_option(self, name, value):
        self._ck(self.L.clb_set_option(self.h, name.encode(), float(value)))

    def get_option(self, name):
        v = C.c_double()
        self._ck(self.L.clb_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    # ---- particles
    def set_particles(self, ids, type, pos, mass, vel=None, q=None, state=None, res_id=None):
        ids, type, pos, mass = _i64(ids), _i32(type), _f64(pos), _f64(mass)
        vel, q, state, res_id = _f64(vel), _f64(q), _i32(state), _i32(res_id)
        self.n = len(ids)
        self._ck(self.L.clb_set_particles(self.h, self.n, _p(ids, c_i64p), _p(type, c_i32p), _p(pos, c_f64p), _p(vel, c_f64p),
                                          _p(mass, c_f64p), _p(q, c_f64p), _p(state, c_i32p), _p(res_id, c_i32p)))

    def num_particles(self):
        return int(self.L.clb_num_particles(self.h))

    _FIELD_SPEC = dict(pos=(3, np.float64), image=(3, np.int32), vel=(3, np.float64), force=(3, np.float64), type=(1, np.int32),
                       state=(1, np.int32), mass=(1, np.float64), q=(1, np.float64), res_id=(1, np.int32))

    def get_particles(self, ids=None, fields=("pos", "vel", "force", "type", "state", "mass", "image", "q", "res_id"), out=None):
        """Per-particle state in ascending-id order (or for `ids`).  `out` may hold caller-owned C-contiguous arrays of the right
        shape and dtype (e.g. pinned host memory): the engine then copies straight into them."""
        n = self.num_particles() if ids is None else len(ids)
        ids_a = _i64(ids)
        res = {}
        for f in fields:
            w, dt = self._FIELD_SPEC[f]
            shape = (n, 3) if w == 3 else (n,)
            buf = out.get(f) if out is not None else None
            if buf is not None:
                if buf.shape != shape or buf.dtype != dt or not buf.flags["C_CONTIGUOUS"]:
                    raise ValueError("out[%r] must be a C-contiguous %s array of shape %s" % (f, np.dtype(dt).name, shape))
            else:
                buf = np.empty(shape, dt)
            res[f] = buf
        out = res
        g = out.get
        self._ck(self.L.clb_get_particles(self.h, n, _p(ids_a, c_i64p), _p(g("pos"), c_f64p), _p(g("image"), c_i32p), _p(g("vel"), c_f64p),
                                          _p(g("force"), c_f64p), _p(g("type"), c_i32p), _p(g("state"), c_i32p), _p(g("mass"), c_f64p),
                                          _p(g("q"), c_f64p), _p(g("res_id"), c_i32p)))
        return out

    FIELDS = dict(type=0, state=1, mass=2, q=3, res_id=4, pos=5, v=6, vel=6)

    def modify_particle(self, pid, field, value):
        v = _f64(np.atleast_1d(value))
        f = self.FIELDS[field] if isinstance(field, str) else int(field)
        self._ck(self.L.clb_modify_particle(self.h, int(pid), f, _p(v, c_f64p)))

    def set_velocities(self, vel):
        vel = _f64(vel)
        self._ck(self.L.clb_set_velocities(self.h, len(vel), _p(vel, c_f64p)))

    def set_positions(self, pos):
        pos = _f64(pos)
        self._ck(self.L.clb_set_positions(self.h, len(pos), _p(pos, c_f64p)))

    # ---- exclusions
    def set_exclusions(self, pairs):
        pairs = _i64(pairs).reshape(-1, 2)
        self._ck(self.L.clb_set_exclusions(self.h, len(pairs), _p(pairs, c_i64p)))

    def get_exclusions(self):
        n = int(self.L.clb_num_exclusions(self.h))
        out = np.zeros((max(n, 1), 2), np.int64)
        m = C.c_int64()
        self._ck(self.L.clb_get_exclusions(self.h, n, _p(out, c_i64p), C.byref(m)))
        return out[:m.value]

    def exclusions_observe(self, lst):
        self._ck(self.L.clb_exclusions_observe(self.h, int(lst)))

    # ---- tables / non-bonded
    def add_table(self, x, e, f, interp=1):
        x, e, f = _f64(x), _f64(e), _f64(f)
        h = C.c_int()
        self._ck(self.L.clb_add_table(self.h, len(x), _p(x, c_f64p), _p(e, c_f64p), _p(f, c_f64p), int(interp), C.byref(h)))
        return h.value

    def add_nonbonded(self, kind):
        h = C.c_int()
        k = NB[kind] if isinstance(kind, str) else int(kind)
        self._ck(self.L.clb_add_nonbonded(self.h, k, C.byref(h)))
        return h.value

    def nb_set_tabulated(self, inter, t1, t2, table, cutoff):
        self._ck(self.L.clb_nb_set_tabulated(self.h, inter, t1, t2, table, float(cutoff)))

    def nb_set_lj(self, inter, t1, t2, eps, sig, cutoff, shift_auto=1):
        self._ck(self.L.clb_nb_set_lj(self.h, inter, t1, t2, float(eps), float(sig), float(cutoff), int(shift_auto)))

    def nb_set_mixed(self, inter, t1, t2, tab1, tab2, mix, conv_type, conv_total, cutoff):
        self._ck(self.L.clb_nb_set_mixed(self.h, inter, t1, t2, tab1, tab2, float(mix), int(conv_type), float(conv_total), float(cutoff)))

    # ---- lists / bonded
    def add_list(self, arity):
        h = C.c_int()
        self._ck(self.L.clb_add_list(self.h, int(arity), C.byref(h)))
        return h.value

    def list_add(self, lst, ids):
        ids = _i64(ids)
        n = 0 if ids.size == 0 else len(ids.reshape(-1, ids.shape[-1]))
        self._ck(self.L.clb_list_add(self.h, int(lst), n, _p(ids, c_i64p)))

    def list_size(self, lst):
        return int(self.L.clb_list_size(self.h, int(lst)))

    def list_get(self, lst, arity):
        n = self.list_size(lst)
        out = np.zeros((max(n, 1), arity), np.int64)
        m = C.c_int64()
        self._ck(self.L.clb_list_get(self.h, int(lst), n, _p(out, c_i64p), C.byref(m)))
        return out[:n]

    def add_bonded(self, lst, typed=0):
        h = C.c_int()
        self._ck(self.L.clb_add_bonded(self.h, int(lst), int(typed), C.byref(h)))
        return h.value

    def bonded_set_potential(self, inter, types, kind, params=(), table=-1):
        t = list(types) + [-1] * (4 - len(types))
        k = POT[kind] if isinstance(kind, str) else int(kind)
        p = _f64(list(params) if len(params) else [0.0])
        self._ck(self.L.clb_bonded_set_potential(self.h, inter, t[0], t[1], t[2], t[3], k, _p(p, c_f64p), len(params), int(table)))

    # ---- observables
    def energy(self, inter):
        v = C.c_double()
        self._ck(self.L.clb_energy(self.h, int(inter), C.byref(v)))
        return v.value

    def kinetics(self):
        out = np.zeros(3)
        self._ck(self.L.clb_kinetics(self.h, _p(out, c_f64p)))
        return out

    def count_type(self, type, state=-1):
        v = C.c_int64()
        self._ck(self.L.clb_count_type(self.h, int(type), int(state), C.byref(v)))
        return v.value

    # ---- integrator
    def set_dt(self, dt):
        self._ck(self.L.clb_set_dt(self.h, float(dt)))

    def set_langevin(self, enabled, kT, gamma, types=()):
        t = _i32(list(types) if len(types) else [0])
        self._ck(self.L.clb_set_langevin(self.h, int(enabled), float(kT), float(gamma), len(types), _p(t, c_i32p)))

    def run(self, n):
        self._ck(self.L.clb_run(self.h, int(n)))

    def set_cap_force(self, cap):
        """integrator.CapForce(system, cap): cap <= 0 switches it off."""
        self._ck(self.L.clb_set_cap_force(self.h, float(cap)))

    def run_continue(self, n):
        """Next chunk of the same integrator.run(n): no run-entry force recalculation / heat-up (clb_run_continue)."""
        self._ck(self.L.clb_run_continue(self.h, int(n)))

    def step(self):
        return int(self.L.clb_step(self.h))

    def decompose(self):
        self._ck(self.L.clb_decompose(self.h))

    def compute_forces(self):
        self._ck(self.L.clb_compute_forces(self.h))

    # ---- reactions
    def reaction_general(self, enabled, interval, nearest, max_per_interval=0):
        self._ck(self.L.clb_reaction_general(self.h, int(enabled), int(interval), int(nearest), int(max_per_interval)))

    def add_reaction(self, type_1, type_2, delta_1, delta_2, min_state_1, max_state_1, min_state_2, max_state_2, rate, cutoff,
                     lst, min_cutoff=0.0, intramolecular=1, intraresidual=1, is_virtual=0, active=1):
        s = ReactionSpec(type_1, type_2, delta_1, delta_2, min_state_1, max_state_1, min_state_2, max_state_2, rate, cutoff,
                         min_cutoff, lst, int(intramolecular), int(intraresidual), int(is_virtual), int(active))
        h = C.c_int()
        self._ck(self.L.clb_add_reaction(self.h, C.byref(s), C.byref(h)))
        return h.value

    def reaction_set_rate(self, r, rate):
        self._ck(self.L.clb_reaction_set_rate(self.h, int(r), float(rate)))

    def reaction_set_active(self, r, active):
        self._ck(self.L.clb_reaction_set_active(self.h, int(r), int(active)))

    def reaction_define_connections(self, r, pairs):
        """RestrictReaction.define_connection for every pair of the connectivity map (reaction_setup.py:115-126)."""
        p = np.ascontiguousarray(np.asarray(pairs, np.int64).reshape(-1, 2))
        self._ck(self.L.clb_reaction_define_connections(self.h, int(r), len(p), _p(p, c_i64p)))

    def reaction_add_change(self, reaction, side, nb_level, old_type, new_type, new_mass=-1.0, new_q=float("nan"),
                            state_mode=0, state_value=0):
        self._ck(self.L.clb_reaction_add_change(self.h, int(reaction), int(side), int(nb_level), int(old_type), int(new_type),
                                                float(new_mass), float(new_q), int(state_mode), int(state_value)))

    def topology_observe(self, lst):
        self._ck(self.L.clb_topology_observe(self.h, int(lst)))

    def topology_register(self, lst, types):
        if len(types) == 3:
            self._ck(self.L.clb_topology_register_triplet(self.h, int(lst), *[int(t) for t in types]))
        else:
            self._ck(self.L.clb_topology_register_quadruplet(self.h, int(lst), *[int(t) for t in types]))

    def topology_initialize(self):
        self._ck(self.L.clb_topology_initialize(self.h))

    def react_now(self):
        v = C.c_int64()
        self._ck(self.L.clb_react_now(self.h, C.byref(v)))
        return v.value

    def reaction_counters(self, n):
        out = np.zeros(max(n, 1), np.int64)
        self._ck(self.L.clb_reaction_counters(self.h, n, _p(out, c_i64p)))
        return out[:n]

    # ---- ATRPActivator
    def atrp_configure(self, num_particles, ratio_activator, ratio_deactivator, delta_catalyst, k_activate, k_deactivate):
        self._ck(self.L.clb_atrp_configure(self.h, int(num_particles), float(ratio_activator), float(ratio_deactivator), float(delta_catalyst),
                                           float(k_activate), float(k_deactivate)))

    def atrp_add_center(self, type_id, state, needs_deactivator, new_type=-1, new_mass=-1.0, new_q=float("nan"), delta_state=0):
        self._ck(self.L.clb_atrp_add_center(self.h, int(type_id), int(state), int(bool(needs_deactivator)), int(new_type), float(new_mass),
                                            float(new_q), int(delta_state)))

    def atrp_now(self):
        """One ATRPActivator pass at the current step: ((activated, deactivated), (ratio_activator, ratio_deactivator))."""
        cnt = np.zeros(2, np.int64); rat = np.zeros(2)
        self._ck(self.L.clb_atrp_now(self.h, _p(cnt, c_i64p), _p(rat, c_f64p)))
        return (int(cnt[0]), int(cnt[1])), (float(rat[0]), float(rat[1]))

    # ---- parity / introspection
    def pairs(self):
        m = C.c_int64()
        self._ck(self.L.clb_get_pairs(self.h, 0, None, C.byref(m)))
        out = np.zeros((max(m.value, 1), 2), np.int64)
        self._ck(self.L.clb_get_pairs(self.h, m.value, _p(out, c_i64p), C.byref(m)))
        return out[:m.value]

    def last_candidates(self):
        m = C.c_int64()
        self._ck(self.L.clb_get_last_candidates(self.h, 0, None, None, C.byref(m)))
        rows = np.zeros((max(m.value, 1), 4), np.int64)
        d2 = np.zeros(max(m.value, 1))
        self._ck(self.L.clb_get_last_candidates(self.h, m.value, _p(rows, c_i64p), _p(d2, c_f64p), C.byref(m)))
        return rows[:m.value], d2[:m.value]

    TIMER_NAMES = ("pair", "bonded", "neighbour", "integrate", "comm", "reaction", "other", "total")
    COUNTER_NAMES = ("steps", "rebuilds", "launches", "list_entries", "reaction_passes", "reaction_events", "ghosts", "interacting_pairs")

    def timers(self):
        t = np.zeros(8); c = np.zeros(8, np.int64)
        self._ck(self.L.clb_timers(self.h, _p(t, c_f64p), _p(c, c_i64p)))
        return dict(zip(self.TIMER_NAMES, t.tolist())), dict(zip(self.COUNTER_NAMES, c.tolist()))

    def reset_timers(self):
        self._ck(self.L.clb_reset_timers(self.h))

    def device_ptr(self, which):
        p = C.c_void_p(); n = C.c_int64()
        self._ck(self.L.clb_device_ptr(self.h, int(which), C.byref(p), C.byref(n)))
        return p.value, n.value

    def stream(self):
        s = C.c_void_p()
        self._ck(self.L.clb_stream(self.h, C.byref(s)))
        return s.value

    # ---- multi-GPU
    def comm_init(self, rank, nranks, nccl_id):
        buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
        self._ck(self.L.clb_comm_init(self.h, int(rank), int(nranks), buf))

    def join(self, group=None):
        """Make this engine one slab of a multi-GPU run: rank 0 creates the NCCL id, torch.distributed broadcasts
        the 128 bytes (torch is plumbing only), every rank calls clb_comm_init.  Replaces the MPI node grid of
        storage.DomainDecomposition (src/start_simulation.py:152-163).  Must precede set_particles."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world == 1:
            return rank, world
        self.comm_init(rank, world, broadcast_nccl_id(group))
        return rank, world

    @staticmethod
    def nccl_unique_id():
        L = _lib.load()
        buf = (C.c_char * 128)()
        rc = L.clb_nccl_unique_id(buf)
        if rc != 0:
            raise EngineError(L.clb_last_error(None).decode())
        return bytes(buf)
