"""ctypes binding of the C ABI declared in include/parm_b200.h.

This is the only way Python reaches the product: every call lands in
parm_b200/libparm_b200.so (hand-written sm_100a CUDA). There is no CPU fallback --
if the library is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libparm_b200.so")

OK, ERR_INVALID, ERR_RUNTIME, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
X, V, A, F, M, ALL = 1, 2, 4, 8, 16, 31
RED_MASS, RED_MOMENTUM, RED_KE, RED_COM, RED_NDOF, RED_COMFORCE = range(6)
(INTEG_VERLET, INTEG_SOL, INTEG_DAMPED, INTEG_SOLHT, INTEG_OVERDAMPED, INTEG_NOSEHOOVER, INTEG_GAUSSIANT, INTEG_GEAR3A,
 INTEG_GEAR4A, INTEG_GEAR5A, INTEG_GEAR6A, INTEG_NLCG) = range(12)
PAIR_LJREPULSE, PAIR_REPULSION, PAIR_LJATTRACTREPULSE, PAIR_LJCUT = range(4)
(PAIR_LJATTRACTCUT, PAIR_LJATTRACTFIXEDREPULSE, PAIR_EISMCLACHLAN, PAIR_LJISH, PAIR_LJATTRACTREPULSESIGS,
 PAIR_REPULSIONDRAG, PAIR_LOISOHERN, PAIR_LOISLIN, PAIR_LOISOHERNMIN, PAIR_LOISLINMIN) = range(4, 14)
WANT_ENERGY, WANT_VIRIAL, WANT_STRESS = 1, 2, 4

dp = C.POINTER(C.c_double)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/parm_b200.h one to one
SIGNATURES = {
    "parm_b200_last_error": (C.c_char_p, []),
    "parm_b200_launch_count": (C.c_uint64, []),
    "parm_b200_version": (C.c_char_p, []),
    "parm_b200_probe_peaks": (C.c_int, [C.c_int, dp, dp]),
    "parm_ctx_create": (C.c_int, [C.c_int, C.c_uint32, C.c_int, vpp]),
    "parm_ctx_destroy": (C.c_int, [vp]),
    "parm_set_box": (C.c_int, [vp, dp]),
    "parm_get_box": (C.c_int, [vp, dp]),
    "parm_box_diff": (C.c_int, [vp, C.c_uint32, dp, dp, dp]),
    "parm_upload_atoms": (C.c_int, [vp, C.c_uint, dp, dp, dp, dp, dp, C.c_size_t, C.c_size_t]),
    "parm_download_atoms": (C.c_int, [vp, C.c_uint, dp, dp, dp, dp, dp, C.c_size_t, C.c_size_t]),
    "parm_host_register": (C.c_int, [vp, C.c_size_t]),
    "parm_host_unregister": (C.c_int, [vp]),
    "parm_sync": (C.c_int, [vp]),
    "parm_snapshot_begin": (C.c_int, [vp, C.c_uint]),
    "parm_snapshot_wait": (C.c_int, [vp, dp, dp]),
    "parm_grid_locs": (C.c_int, [vp, u32p, u32p]),
    "parm_get_stream": (C.c_int, [vp, vpp]),
    "parm_profile_enable": (C.c_int, [vp, C.c_int]),
    "parm_profile_read": (C.c_int, [vp, dp, u64p]),
    "parm_reduce": (C.c_int, [vp, C.c_int, dp, dp]),
    "parm_scale_velocities": (C.c_int, [vp, C.c_double]),
    "parm_add_velocity": (C.c_int, [vp, dp]),
    "parm_reset_forces": (C.c_int, [vp]),
    "parm_nlist_create": (C.c_int, [vp, C.c_double, vpp]),
    "parm_nlist_destroy": (C.c_int, [vp]),
    "parm_nlist_set_diameters": (C.c_int, [vp, dp]),
    "parm_nlist_update": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int)]),
    "parm_nlist_which": (C.c_int, [vp, u32p]),
    "parm_nlist_numpairs": (C.c_int, [vp, u64p]),
    "parm_nlist_download_pairs": (C.c_int, [vp, u32p, u32p, C.c_uint64]),
    "parm_nlist_stats": (C.c_int, [vp, dp, u32p]),
    "parm_nlist_tile_stats": (C.c_int, [vp, C.POINTER(C.c_int), u32p, u32p, u32p]),
    "parm_nlist_ignore": (C.c_int, [vp, u32p, u32p, C.c_uint64]),
    "parm_nlist_ignore_size": (C.c_int, [vp, u64p]),
    "parm_inter_create": (C.c_int, [vp, vp, C.c_int, vpp]),
    "parm_inter_destroy": (C.c_int, [vp]),
    "parm_inter_set_params": (C.c_int, [vp, dp, u32p, dp, C.c_int, u8p, C.c_int]),
    "parm_inter_set_params_ex": (C.c_int, [vp, dp, C.c_int, u32p, dp, dp, C.c_int, u8p, C.c_int]),
    "parm_inter_set_forces": (C.c_int, [vp, C.c_uint, dp]),
    "parm_inter_energy": (C.c_int, [vp, dp]),
    "parm_inter_pressure": (C.c_int, [vp, dp]),
    "parm_inter_stress": (C.c_int, [vp, dp]),
    "parm_inter_contacts": (C.c_int, [vp, u64p, u64p]),
    "parm_verlet_create": (C.c_int, [vp, C.c_double, vpp]),
    "parm_sol_create": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, C.c_uint64, vpp]),
    "parm_integ_create": (C.c_int, [vp, C.c_int, dp, C.c_int, C.c_uint64, vpp]),
    "parm_nlcg_create": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint, C.c_double, vpp]),
    "parm_nlcg_set": (C.c_int, [vp, C.c_int, C.c_double]),
    "parm_nlcg_get": (C.c_int, [vp, dp]),
    "parm_nlcg_set_forces": (C.c_int, [vp, C.c_int, C.c_int]),
    "parm_nlcg_reset": (C.c_int, [vp]),
    "parm_nlcg_descend": (C.c_int, [vp]),
    "parm_nlcg_reduce": (C.c_int, [vp, C.c_int, dp]),
    "parm_rsq_create": (C.c_int, [vp, u64p, C.c_int, C.c_int, vpp]),
    "parm_isf_create": (C.c_int, [vp, dp, C.c_int, u64p, C.c_int, C.c_int, vpp]),
    "parm_energy_tracker_create": (C.c_int, [vp, vpp, C.c_int, C.c_uint, vpp]),
    "parm_tracker_destroy": (C.c_int, [vp]),
    "parm_tracker_update": (C.c_int, [vp]),
    "parm_tracker_reset": (C.c_int, [vp]),
    "parm_tracker_counts": (C.c_int, [vp, u64p, C.c_int]),
    "parm_rsq_read": (C.c_int, [vp, C.c_int, dp, dp, dp]),
    "parm_isf_read": (C.c_int, [vp, C.c_int, dp]),
    "parm_energy_tracker_read": (C.c_int, [vp, dp]),
    "parm_energy_tracker_set_u0": (C.c_int, [vp, C.c_int, C.c_double]),
    "parm_integ_add_stat_tracker": (C.c_int, [vp, vp]),
    "parm_integ_register_stat_tracker": (C.c_int, [vp, vp]),
    "parm_integ_get_scalars": (C.c_int, [vp, dp]),
    "parm_integ_reset_bath": (C.c_int, [vp]),
    "parm_integ_set_param": (C.c_int, [vp, C.c_int, C.c_double]),
    "parm_integ_destroy": (C.c_int, [vp]),
    "parm_integ_add_interaction": (C.c_int, [vp, vp]),
    "parm_integ_add_tracker": (C.c_int, [vp, vp]),
    "parm_integ_register_interaction": (C.c_int, [vp, vp]),
    "parm_integ_register_tracker": (C.c_int, [vp, vp]),
    "parm_integ_initialize": (C.c_int, [vp]),
    "parm_integ_set_dt": (C.c_int, [vp, C.c_double]),
    "parm_integ_set_temperature": (C.c_int, [vp, C.c_double, C.c_double]),
    "parm_integ_set_forces": (C.c_int, [vp, C.c_int]),
    "parm_integ_timestep": (C.c_int, [vp, C.c_int]),
    "parm_integ_update_trackers": (C.c_int, [vp]),
    "parm_integ_potential_energy": (C.c_int, [vp, dp]),
    "parm_integ_virial": (C.c_int, [vp, dp]),
    "parm_integ_inject_noise": (C.c_int, [vp, dp, C.c_size_t]),
    "parm_integ_get_sol_constants": (C.c_int, [vp, dp]),
    "parm_integ_stats": (C.c_int, [vp, u64p, u64p, u64p]),
    "parm_nccl_unique_id": (C.c_int, [vp]),
    "parm_ctx_create_sharded": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, vp, vpp]),
    "parm_shard_set_atoms": (C.c_int, [vp, C.c_uint32, u32p, dp, dp, dp, dp, dp]),
    "parm_shard_get_atoms": (C.c_int, [vp, C.c_uint32, u32p, u32p, dp, dp, dp, dp, dp]),
    "parm_shard_put_atoms": (C.c_int, [vp, C.c_uint32, dp, dp, dp, dp]),
    "parm_shard_info": (C.c_int, [vp, u32p]),
    "parm_shard_rebuild_stats": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
}


class ParmError(RuntimeError):
    pass


class ParmInvalid(ValueError):
    """std::invalid_argument in the reference (SWIG maps it to ValueError, sim.i:139-157)."""


class ParmUnsupported(NotImplementedError):
    pass


_lib = None


def lib():
    """Load libparm_b200.so (once). Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ParmError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(or make -C parm_b200/csrc). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc == OK:
        return
    msg = lib().parm_b200_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ParmInvalid(msg)
    if rc == ERR_UNSUPPORTED:
        raise ParmUnsupported(msg)
    raise ParmError(msg)


def call(name, *args):
    check(getattr(lib(), name)(*args))
