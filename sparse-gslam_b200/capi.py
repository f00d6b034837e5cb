"""ctypes mirror of include/sgb_capi.h (the C ABI a g2o plugin binds to). No torch types, plain pointers."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGB_LIB") or os.path.join(HERE, "libsgb.so")  # SGB_LIB: a variant build (tuning runs)

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_NOT_INITIALIZED, ERR_UNSUPPORTED, ERR_SOLVE_FAILED, ERR_COMM = range(8)
RESULT_TERMINATE, RESULT_OK, RESULT_FAIL = 2, 1, -1
ALGO_LM, ALGO_GN = 0, 1
JAC_G2O_NUMERIC, JAC_ANALYTIC = 0, 1

EXPORTS = [
    "sgb_version", "sgb_device_count", "sgb_default_options", "sgb_create", "sgb_destroy", "sgb_last_error",
    "sgb_set_graph", "sgb_get_structure_info", "sgb_get_structure", "sgb_linearize", "sgb_solve_once", "sgb_optimize",
    "sgb_step", "sgb_get_estimates", "sgb_set_estimates", "sgb_push", "sgb_pop", "sgb_discard_top", "sgb_chi2",
    "sgb_get_timings", "sgb_optimize_resident", "sgb_set_graph_partitioned", "sgb_comm_get_handle", "sgb_comm_connect",
    "sgb_get_partition_info", "sgb_optimize_batch", "sgb_optimize_batch_resident", "sgb_g2o_load", "sgb_g2o_view",
    "sgb_g2o_free", "sgb_g2o_save",
    "sgb_compute_marginals", "sgb_update_graph",
    "sgb_linear_set_pattern", "sgb_linear_solve", "sgb_set_graph_device", "sgb_pg_create", "sgb_pg_destroy", "sgb_pg_last_error", "sgb_pg_reset",
    "sgb_pg_append_from_lm", "sgb_pg_append_from_host", "sgb_pg_add_closure", "sgb_pg_optimize",
    "sgb_pg_prune_closures", "sgb_pg_get_info", "sgb_pg_download",
    "sgb_odom_information", "sgb_scan_point_covariances", "sgb_line_fit_information", "sgb_frontend_last_error",
]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("jacobian_mode", C.c_int32), ("pcg_tolerance", C.c_double),
                ("pcg_max_iters", C.c_int32), ("verbose", C.c_int32), ("lm_tau", C.c_double),
                ("lm_user_lambda", C.c_double), ("lm_max_trials", C.c_int32), ("incremental", C.c_int32),
                ("coarse_nodes", C.c_int32)]


class GraphSoA(C.Structure):
    _fields_ = [
        ("n_poses", C.c_int32), ("pose_id", C.c_void_p), ("pose_est", C.c_void_p), ("pose_fixed", C.c_void_p),
        ("n_landmarks", C.c_int32), ("lm_id", C.c_void_p), ("lm_est", C.c_void_p), ("lm_fixed", C.c_void_p),
        ("n_pp", C.c_int32), ("pp_i", C.c_void_p), ("pp_j", C.c_void_p), ("pp_z", C.c_void_p), ("pp_info", C.c_void_p),
        ("pp_phi", C.c_void_p), ("pp_seq", C.c_void_p),
        ("n_pl", C.c_int32), ("pl_pose", C.c_void_p), ("pl_lm", C.c_void_p), ("pl_z", C.c_void_p),
        ("pl_info", C.c_void_p), ("pl_seq", C.c_void_p),
    ]


class GraphDelta(C.Structure):
    """sgb_graph_delta: the new vertices / edges of sgb_update_graph"""
    _fields_ = [("n_new_poses", C.c_int32), ("pose_id", C.c_void_p), ("pose_est", C.c_void_p), ("pose_fixed", C.c_void_p),
                ("n_new_landmarks", C.c_int32), ("lm_id", C.c_void_p), ("lm_est", C.c_void_p), ("lm_fixed", C.c_void_p),
                ("n_new_pp", C.c_int32), ("pp_i", C.c_void_p), ("pp_j", C.c_void_p), ("pp_z", C.c_void_p),
                ("pp_info", C.c_void_p), ("pp_phi", C.c_void_p), ("pp_seq", C.c_void_p),
                ("n_new_pl", C.c_int32), ("pl_pose", C.c_void_p), ("pl_lm", C.c_void_p), ("pl_z", C.c_void_p),
                ("pl_info", C.c_void_p), ("pl_seq", C.c_void_p)]


class IterStat(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("trials", C.c_int32), ("result", C.c_int32), ("pcg_iters", C.c_int32),
                ("chi2", C.c_double), ("lambda_", C.c_double), ("rho", C.c_double), ("chi2_before", C.c_double),
                ("pcg_residual", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class StructureInfo(C.Structure):
    _fields_ = [("n_free", C.c_int32), ("n_free_poses", C.c_int32), ("n_free_landmarks", C.c_int32),
                ("n_blocks", C.c_int32), ("scalar_dim", C.c_int32), ("n_active_pp", C.c_int32),
                ("n_active_pl", C.c_int32), ("coarse_nodes", C.c_int32), ("block_values", C.c_int64)]


class Timings(C.Structure):
    _fields_ = [("linearize_ms", C.c_double), ("setup_ms", C.c_double), ("pcg_ms", C.c_double),
                ("update_ms", C.c_double), ("total_ms", C.c_double), ("pcg_iters", C.c_int64), ("trials", C.c_int64),
                ("linearizations", C.c_int64), ("kernel_launches", C.c_int64), ("pcg_phase_ms", C.c_double * 4)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["pcg_phase_ms"] = list(self.pcg_phase_ms)
        return d


class PartitionInfo(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("n_poses", C.c_int32), ("n_landmarks", C.c_int32),
                ("n_pp", C.c_int32), ("n_pl", C.c_int32), ("n_pp_owned", C.c_int32), ("n_pl_owned", C.c_int32),
                ("halo_pose_gathers", C.c_int64), ("halo_landmark_gathers", C.c_int64), ("hpp_entries", C.c_int64),
                ("hpl_entries", C.c_int64), ("hlp_entries", C.c_int64), ("hpp_blocks", C.c_int64),
                ("hpl_blocks", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BlockMatrix(C.Structure):
    _fields_ = [("n_block_cols", C.c_int32), ("block_dim", C.c_void_p), ("col_ptr", C.c_void_p), ("row_idx", C.c_void_p)]


class DeviceValues(C.Structure):
    _fields_ = [("pose_est", C.c_void_p), ("lm_est", C.c_void_p), ("pp_z", C.c_void_p), ("pp_info", C.c_void_p),
                ("pp_phi", C.c_void_p), ("pp_slot", C.c_void_p), ("pl_z", C.c_void_p), ("pl_info", C.c_void_p),
                ("pl_slot", C.c_void_p), ("has_robust", C.c_int32), ("reserved", C.c_int32)]


class PgInfo(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_edges", C.c_int32), ("n_closures", C.c_int32),
                ("n_active_closures", C.c_int32), ("last_edit_ms", C.c_double), ("kernel_launches", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load() -> C.CDLL:
    """Load libsgb.so. Raises loudly when it has not been built: there is no Python/CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m sparse_gslam_b200.build` "
                           "(the backend is CUDA-only; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.sgb_version.restype = C.c_char_p
    L.sgb_device_count.restype = C.c_int32
    L.sgb_default_options.argtypes = [C.POINTER(Options)]
    L.sgb_create.argtypes = [C.POINTER(Options), C.POINTER(vp)]
    L.sgb_destroy.argtypes = [vp]
    L.sgb_last_error.argtypes = [vp]
    L.sgb_last_error.restype = C.c_char_p
    L.sgb_set_graph.argtypes = [vp, C.POINTER(GraphSoA)]
    L.sgb_set_graph_partitioned.argtypes = [vp, C.POINTER(GraphSoA), C.c_int32, C.c_int32]
    L.sgb_comm_get_handle.argtypes = [vp, vp]
    L.sgb_comm_connect.argtypes = [vp, vp, C.c_int32]
    L.sgb_get_partition_info.argtypes = [vp, C.POINTER(PartitionInfo)]
    L.sgb_get_structure_info.argtypes = [vp, C.POINTER(StructureInfo)]
    L.sgb_get_structure.argtypes = [vp] + [vp] * 9
    L.sgb_linearize.argtypes = [vp, vp, vp, vp]
    L.sgb_solve_once.argtypes = [vp, C.c_double, vp, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.sgb_optimize.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), vp]
    L.sgb_optimize_resident.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_int32), vp]
    L.sgb_step.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), vp]
    L.sgb_get_estimates.argtypes = [vp, vp, vp]
    L.sgb_set_estimates.argtypes = [vp, vp, vp]
    for f in ("sgb_push", "sgb_pop", "sgb_discard_top"):
        getattr(L, f).argtypes = [vp]
    L.sgb_chi2.argtypes = [vp, vp]
    L.sgb_get_timings.argtypes = [vp, C.POINTER(Timings)]
    for f in ("sgb_optimize_batch", "sgb_optimize_batch_resident"):
        getattr(L, f).argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp]
    L.sgb_g2o_load.argtypes = [C.c_char_p, C.POINTER(vp), C.c_char_p, C.c_int32]
    L.sgb_g2o_view.argtypes = [vp, C.POINTER(GraphSoA)]
    L.sgb_g2o_free.argtypes = [vp]
    L.sgb_g2o_save.argtypes = [C.c_char_p, C.POINTER(GraphSoA)]
    L.sgb_linear_set_pattern.argtypes = [vp, C.POINTER(BlockMatrix)]
    L.sgb_linear_solve.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.sgb_compute_marginals.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.sgb_update_graph.argtypes = [vp, C.POINTER(GraphDelta)]
    L.sgb_set_graph_device.argtypes = [vp, C.POINTER(GraphSoA), C.POINTER(DeviceValues)]
    L.sgb_pg_create.argtypes = [C.c_int32, C.POINTER(vp)]
    L.sgb_pg_destroy.argtypes = [vp]
    L.sgb_pg_last_error.argtypes = [vp]
    L.sgb_pg_last_error.restype = C.c_char_p
    L.sgb_pg_reset.argtypes = [vp, C.c_int32, vp]
    L.sgb_pg_append_from_lm.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, vp]
    L.sgb_pg_append_from_host.argtypes = [vp, vp, C.c_int32, vp, vp]
    L.sgb_pg_add_closure.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, C.c_double, C.POINTER(C.c_int32)]
    L.sgb_pg_optimize.argtypes = [vp, vp, C.c_int32, C.c_int32, C.POINTER(C.c_int32), vp]
    L.sgb_pg_prune_closures.argtypes = [vp, C.c_double, C.POINTER(C.c_int32), vp, vp]
    L.sgb_pg_get_info.argtypes = [vp, C.POINTER(PgInfo)]
    L.sgb_pg_download.argtypes = [vp] + [vp] * 9
    L.sgb_odom_information.argtypes = [C.c_int32, vp, vp, C.c_int32, C.c_double, C.c_double, C.c_double, vp, vp, vp,
                                       C.POINTER(C.c_double)]
    L.sgb_scan_point_covariances.argtypes = [C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_float, C.c_float,
                                             C.c_float, C.c_float, vp, vp, vp, C.POINTER(C.c_double)]
    L.sgb_line_fit_information.argtypes = [C.c_int32, vp, vp, vp, C.c_int32, vp, vp, vp, C.POINTER(C.c_double)]
    L.sgb_frontend_last_error.restype = C.c_char_p
    _lib = L
    return L
