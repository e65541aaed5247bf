"""ctypes binding of the C ABI in include/chemps2_b200.h (libchemps2_b200.so, built in-tree by `make`).

The library is the product: it is loaded eagerly and a missing .so is an ImportError — there is no Python/CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchemps2_b200.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: run `make` (or __graft_entry__.build()) first; there is no fallback path")
lib = C.CDLL(LIB_PATH)

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_int64)
vp = C.c_void_p


class FlatTerm(C.Structure):
    _fields_ = [("dst", C.c_int32), ("src", C.c_int32), ("a_rows", C.c_int32), ("a_cols", C.c_int32), ("b_rows", C.c_int32),
                ("b_cols", C.c_int32), ("a_space", C.c_int8), ("a_trans", C.c_int8), ("b_space", C.c_int8), ("b_trans", C.c_int8),
                ("owner", C.c_int32), ("a_off", C.c_int64), ("b_off", C.c_int64), ("factor", C.c_double)]


class FlatPresum(C.Structure):
    _fields_ = [("dst_off", C.c_int64), ("src_off", C.c_int64), ("size", C.c_int64), ("space", C.c_int32), ("coef", C.c_double)]


# every symbol include/chemps2_b200.h declares: (name, restype, argtypes)
SIGNATURES = [
    ("b2_last_error", C.c_char_p, []),
    ("b2_version", C.c_char_p, []),
    ("b2_ctx_create", C.c_int, [C.c_int, C.POINTER(vp)]),
    ("b2_ctx_destroy", None, [vp]),
    ("b2_ctx_device", C.c_int, [vp]),
    ("b2_ctx_set_stream", C.c_int, [vp, vp]),
    ("b2_ctx_stream", vp, [vp]),
    ("b2_opset_fill_hash", C.c_int, [vp, C.c_uint64, C.c_double]),
    ("b2_hash_fill", C.c_int, [c_dp, C.c_int64, C.c_uint64, C.c_uint64, C.c_double]),
    ("b2_problem_set", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_dp, C.c_double]),
    ("b2_problem_set_integrals", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_dp, c_dp, C.c_double]),
    ("b2_problem_update_mx", C.c_int, [vp, c_dp]),
    ("b2_problem_mx", C.c_int, [vp, c_dp]),
    ("b2_wigner6j", C.c_double, [C.c_int] * 6),
    ("b2_wigner9j", C.c_double, [C.c_int] * 9),
    ("b2_small_symmetric_eig", C.c_int, [C.c_int, c_dp, c_dp, c_dp]),
    ("b2_bk_init", C.c_int, [vp, C.c_int]),
    ("b2_bk_set_dim", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("b2_bk_dim", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("b2_bk_fcidim", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("b2_bk_nmin", C.c_int, [vp, C.c_int]),
    ("b2_bk_nmax", C.c_int, [vp, C.c_int]),
    ("b2_bk_twosmin", C.c_int, [vp, C.c_int, C.c_int]),
    ("b2_bk_twosmax", C.c_int, [vp, C.c_int, C.c_int]),
    ("b2_tensor_t_size", C.c_int64, [vp, C.c_int]),
    ("b2_sobject_size", C.c_int64, [vp, C.c_int]),
    ("b2_sobject_nkappa", C.c_int, [vp, C.c_int]),
    ("b2_sobject_table", C.c_int, [vp, C.c_int, c_ip, c_lp]),
    ("b2_opset_create", C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    ("b2_opset_create_correlation", C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    ("b2_opset_offload", C.c_int, [vp]),
    ("b2_opset_reload", C.c_int, [vp]),
    ("b2_opset_offload_file", C.c_int, [vp, C.c_char_p]),
    ("b2_dmrg_set_spill_dir", C.c_int, [vp, C.c_char_p]),
    ("b2_opset_resident", C.c_int, [vp]),
    ("b2_dmrg_set_spill", C.c_int, [vp, C.c_int]),
    ("b2_opset_destroy", None, [vp]),
    ("b2_opset_count", C.c_int, [vp]),
    ("b2_opset_info", C.c_int, [vp, C.c_int, c_ip, c_ip, c_ip, c_lp]),
    ("b2_opset_find", C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
    ("b2_opset_upload", C.c_int, [vp, C.c_int, c_dp]),
    ("b2_opset_download", C.c_int, [vp, C.c_int, c_dp]),
    ("b2_opset_clear", C.c_int, [vp]),
    ("b2_opset_host_arena", c_dp, [vp]),
    ("b2_opset_arena_size", C.c_int64, [vp]),
    ("b2_heff_create", C.c_int, [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.POINTER(vp)]),
    ("b2_heff_destroy", None, [vp]),
    ("b2_heff_veclength", C.c_int64, [vp]),
    ("b2_heff_apply", C.c_int, [vp, c_dp, c_dp]),
    ("b2_heff_apply_device", C.c_int, [vp, vp, vp]),
    ("b2_heff_diag", C.c_int, [vp, c_dp]),
    ("b2_heff_stats", C.c_int, [vp, c_dp]),
    ("b2_heff_diag_device", C.c_int, [vp, vp]),
    ("b2_heff_solve", C.c_int, [vp, c_dp, C.c_double, c_dp, c_ip]),
    ("b2_heff_solve_device", C.c_int, [vp, vp, C.c_double, c_dp, c_ip]),
    ("b2_heff_set_allreduce", C.c_int, [vp, vp, vp]),
    ("b2_heff_set_excitations", C.c_int, [vp, C.c_int, C.POINTER(c_dp)]),
    ("b2_heff_last_kernel_seconds", C.c_double, [vp]),
    ("b2_heff_num_terms", C.c_int64, [vp]),
    ("b2_heff_export_terms", C.c_int, [vp, C.POINTER(FlatTerm)]),
    ("b2_heff_num_presum_parts", C.c_int64, [vp]),
    ("b2_heff_presum_size", C.c_int64, [vp]),
    ("b2_heff_export_presums", C.c_int, [vp, C.POINTER(FlatPresum)]),
    ("b2_probe_fp64", C.c_int, [vp, C.c_int, c_dp]),
    ("b2_twodm_fill_site", C.c_int, [vp, C.c_int, c_dp, vp, vp, c_dp, c_dp]),
    ("b2_twodm_create", C.c_int, [vp, C.c_int, vp, vp, C.POINTER(vp)]),
    ("b2_twodm_destroy", None, [vp]),
    ("b2_twodm_run", C.c_int, [vp, c_dp, c_dp, c_dp]),
    ("b2_twodm_worklists", C.c_int, [vp, vp]),
    ("b2_twodm_m_size", C.c_int64, [vp]),
    ("b2_twodm_num_groups", C.c_int, [vp]),
    ("b2_twodm_group_info", C.c_int, [vp, C.c_int, c_ip, c_lp, c_lp, c_lp, c_ip, c_ip, c_ip, C.c_int]),
    ("b2_twodm_d1_scale", C.c_int, [vp, c_dp, C.c_int]),
    ("b2_twodm_scatter", C.c_int, [vp, C.POINTER(c_dp), C.c_double, c_dp, c_dp]),
    ("b2_svd_batch", C.c_int, [vp, C.c_int, c_ip, c_ip, C.POINTER(c_dp), C.POINTER(c_dp), C.POINTER(c_dp), C.POINTER(c_dp)]),
    ("b2_sobject_split", C.c_int, [vp, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(vp), c_dp]),
    ("b2_split_size", C.c_int64, [vp, C.c_int]),
    ("b2_split_get", C.c_int, [vp, C.c_int, c_dp]),
    ("b2_split_destroy", None, [vp]),
    ("b2_heff_diag_lists", C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(vp), C.POINTER(C.c_int64)]),
    ("b2_join_create", C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    ("b2_join_destroy", None, [vp]),
    ("b2_join_run", C.c_int, [vp, c_dp, c_dp, c_dp]),
    ("b2_join_worklists", C.c_int, [vp, vp]),
    ("b2_heff_worklists", C.c_int, [vp, vp]),
    ("b2_ctx_set_option", C.c_int, [vp, C.c_char_p, C.c_double]),
    ("b2_dmrg_create", C.c_int, [vp, C.POINTER(vp)]),
    ("b2_dmrg_destroy", None, [vp]),
    ("b2_dmrg_mps_size", C.c_int64, [vp, C.c_int]),
    ("b2_dmrg_set_mps", C.c_int, [vp, C.c_int, c_dp]),
    ("b2_dmrg_get_mps", C.c_int, [vp, C.c_int, c_dp]),
    ("b2_dmrg_random_mps", C.c_int, [vp, C.c_uint64]),
    ("b2_dmrg_srand", C.c_int, [vp, C.c_uint64]),
    ("b2_davidson_create", C.c_int, [vp, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(vp)]),
    ("b2_davidson_destroy", None, [vp]),
    ("b2_davidson_fetch", C.c_int, [vp, C.POINTER(C.c_char), C.POINTER(vp), C.POINTER(vp)]),
    ("b2_davidson_num_multiplications", C.c_int, [vp]),
    ("b2_davidson_eigenvalue", C.c_double, [vp]),
    ("b2_rand_stream", C.c_int, [C.c_uint64, C.c_int, c_ip]),
    ("b2_dmrg_opset", vp, [vp, C.c_int, C.c_int]),
    ("b2_dmrg_set_opset", C.c_int, [vp, C.c_int, C.c_int, vp]),
    ("b2_dmrg_update", C.c_int, [vp, C.c_int, C.c_int]),
    ("b2_dmrg_solve_site", C.c_int, [vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_ip]),
    ("b2_dmrg_sweep", C.c_int, [vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, c_dp, c_dp]),
    ("b2_dmrg_sweep_info", C.c_int, [vp, c_dp]),
    ("b2_update_create", C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.POINTER(vp)]),
    ("b2_update_create_sharded", C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.POINTER(vp)]),
    ("b2_update_set_allreduce", C.c_int, [vp, vp, vp]),
    ("b2_dmrg_presolve", C.c_int, [vp]),
    ("b2_dmrg_set_plan_cache", C.c_int, [vp, C.c_int]),
    ("b2_dmrg_plan_cache_stats", C.c_int, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    ("b2_dmrg_set_plan_prefetch", C.c_int, [vp, C.c_int]),
    ("b2_dmrg_plan_prefetched", C.c_longlong, [vp]),
    ("b2_dmrg_save_mps", C.c_int, [vp, C.c_char_p, C.c_int]),
    ("b2_dmrg_load_mps", C.c_int, [vp, C.c_char_p, c_ip]),
    ("b2_dmrg_solve", C.c_int, [vp, C.c_int, c_ip, c_dp, c_ip, c_dp, c_dp, c_dp]),
    ("b2_dmrg_new_excitation", C.c_int, [vp, C.c_double, C.c_int, C.c_uint64]),
    ("b2_dmrg_num_lower_states", C.c_int, [vp]),
    ("b2_dmrg_calc_2rdm", C.c_int, [vp, c_dp, c_dp]),
    ("b2_dmrg_calc_correlations", C.c_int, [vp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    ("b2_corr_fill_site", C.c_int, [vp, C.c_int, c_dp, vp, c_dp, c_dp, c_dp, c_dp]),
    ("b2_dmrg_set_world", C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    ("b2_dmrg_timers", C.c_int, [vp, c_dp, C.c_int]),
    ("b2_update_destroy", None, [vp]),
    ("b2_update_run", C.c_int, [vp, c_dp]),
    ("b2_update_run_device", C.c_int, [vp, vp]),
    ("b2_update_stats", C.c_int, [vp, c_dp]),
    ("b2_update_worklists", C.c_int, [vp, C.c_int, vp]),
    ("b2_update_num_presum_parts", C.c_int64, [vp]),
    ("b2_update_num_mix_flat", C.c_int64, [vp]),
    ("b2_update_export_mix_flat", C.c_int, [vp, C.POINTER(FlatPresum)]),
    ("b2_update_presum_size", C.c_int64, [vp]),
    ("b2_update_export_presums", C.c_int, [vp, C.POINTER(FlatPresum)]),
]
for _name, _res, _args in SIGNATURES:
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


class Worklists(C.Structure):
    _fields_ = [("items1", vp), ("items2", vp), ("tiles1", vp * 4), ("tiles2", vp * 4), ("reduces", vp), ("waves", vp),
                ("n_items1", C.c_int64), ("n_items2", C.c_int64), ("n_tiles1", C.c_int64 * 4), ("n_tiles2", C.c_int64 * 4),
                ("n_reduces", C.c_int64), ("n_waves", C.c_int64), ("work_size", C.c_int64), ("part_size", C.c_int64)]


lib.b2_heff_worklists.restype = C.c_int
lib.b2_heff_worklists.argtypes = [vp, C.POINTER(Worklists)]


class B2Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise B2Error(f"chemps2_b200 error {rc}: {lib.b2_last_error().decode()}")
