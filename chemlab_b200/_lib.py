"""ctypes binding of include/chemlab_b200.h.  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "lib", "libchemlab_b200.so")

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)


class ReactionSpec(C.Structure):
    _fields_ = [("type_1", C.c_int32), ("type_2", C.c_int32), ("delta_1", C.c_int32), ("delta_2", C.c_int32),
                ("min_state_1", C.c_int32), ("max_state_1", C.c_int32), ("min_state_2", C.c_int32), ("max_state_2", C.c_int32),
                ("rate", C.c_double), ("cutoff", C.c_double), ("min_cutoff", C.c_double),
                ("list", C.c_int32), ("intramolecular", C.c_int32), ("intraresidual", C.c_int32),
                ("is_virtual", C.c_int32), ("active", C.c_int32)]


# name -> (restype, argtypes); every symbol declared in include/chemlab_b200.h
SIGNATURES = {
    "clb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_f64p, C.c_double, C.c_double, C.c_uint64]),
    "clb_destroy": (None, [C.c_void_p]),
    "clb_last_error": (C.c_char_p, [C.c_void_p]),
    "clb_abi_version": (C.c_int, []),
    "clb_trim_cache": (C.c_int, []),
    "clb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "clb_get_option": (C.c_int, [C.c_void_p, C.c_char_p, c_f64p]),
    "clb_set_particles": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, c_i32p, c_f64p, c_f64p, c_f64p, c_f64p, c_i32p, c_i32p]),
    "clb_num_particles": (C.c_int64, [C.c_void_p]),
    "clb_get_particles": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, c_f64p, c_i32p, c_f64p, c_f64p, c_i32p, c_i32p, c_f64p, c_f64p, c_i32p]),
    "clb_modify_particle": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_f64p]),
    "clb_set_velocities": (C.c_int, [C.c_void_p, C.c_int64, c_f64p]),
    "clb_set_positions": (C.c_int, [C.c_void_p, C.c_int64, c_f64p]),
    "clb_set_exclusions": (C.c_int, [C.c_void_p, C.c_int64, c_i64p]),
    "clb_num_exclusions": (C.c_int64, [C.c_void_p]),
    "clb_get_exclusions": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, c_i64p]),
    "clb_exclusions_observe": (C.c_int, [C.c_void_p, C.c_int]),
    "clb_add_table": (C.c_int, [C.c_void_p, C.c_int64, c_f64p, c_f64p, c_f64p, C.c_int, C.POINTER(C.c_int)]),
    "clb_add_nonbonded": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "clb_nb_set_tabulated": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "clb_nb_set_lj": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]),
    "clb_nb_set_mixed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]),
    "clb_add_list": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "clb_list_add": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i64p]),
    "clb_list_size": (C.c_int64, [C.c_void_p, C.c_int]),
    "clb_list_get": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i64p, c_i64p]),
    "clb_add_bonded": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "clb_bonded_set_potential": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f64p, C.c_int, C.c_int]),
    "clb_energy": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "clb_kinetics": (C.c_int, [C.c_void_p, c_f64p]),
    "clb_count_type": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_i64p]),
    "clb_set_dt": (C.c_int, [C.c_void_p, C.c_double]),
    "clb_set_langevin": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, c_i32p]),
    "clb_set_cap_force": (C.c_int, [C.c_void_p, C.c_double]),
    "clb_run": (C.c_int, [C.c_void_p, C.c_int64]),
    "clb_run_continue": (C.c_int, [C.c_void_p, C.c_int64]),
    "clb_step": (C.c_int64, [C.c_void_p]),
    "clb_decompose": (C.c_int, [C.c_void_p]),
    "clb_compute_forces": (C.c_int, [C.c_void_p]),
    "clb_reaction_general": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "clb_add_reaction": (C.c_int, [C.c_void_p, C.POINTER(ReactionSpec), C.POINTER(C.c_int)]),
    "clb_reaction_set_rate": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "clb_reaction_set_active": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "clb_reaction_define_connections": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i64p]),
    "clb_reaction_add_change": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]),
    "clb_topology_observe": (C.c_int, [C.c_void_p, C.c_int]),
    "clb_topology_register_triplet": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "clb_topology_register_quadruplet": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "clb_topology_initialize": (C.c_int, [C.c_void_p]),
    "clb_react_now": (C.c_int, [C.c_void_p, c_i64p]),
    "clb_reaction_counters": (C.c_int, [C.c_void_p, C.c_int, c_i64p]),
    "clb_atrp_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "clb_atrp_add_center": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]),
    "clb_atrp_now": (C.c_int, [C.c_void_p, c_i64p, c_f64p]),
    "clb_get_pairs": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, c_i64p]),
    "clb_get_last_candidates": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, c_f64p, c_i64p]),
    "clb_timers": (C.c_int, [C.c_void_p, c_f64p, c_i64p]),
    "clb_reset_timers": (C.c_int, [C.c_void_p]),
    "clb_device_ptr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), c_i64p]),
    "clb_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "clb_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "clb_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load():
    """dlopen the engine.  No fallback of any kind: a missing library is an ImportError."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise ImportError("chemlab_b200: %s is missing -- build it with `python -m chemlab_b200.build` "
                          "(or __graft_entry__.build()); there is no CPU fallback" % SO)
    L = C.CDLL(SO, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
