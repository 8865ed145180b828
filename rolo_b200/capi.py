"""ctypes loader for the C-ABI shared library (include/rolo_gicp.h).

The library is the product: hand-written sm_100a CUDA behind `extern "C"` entry points. This module
only loads it and declares prototypes; it never falls back to anything else. If the .so is missing
it is built in-tree with nvcc (rolo_b200/build.py); if that fails, importing raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "rolo_gicp.h")

RG_OK, RG_ERR_INVALID, RG_ERR_CUDA, RG_ERR_NO_DEVICE, RG_ERR_UNSUPPORTED, RG_ERR_NOT_CONVERGED = 0, -1, -2, -3, -4, -5
REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS, REG_PLANE_S = range(6)
DIRECT27, DIRECT7, DIRECT1, DIRECT_RADIUS = range(4)
ACC_ADDITIVE, ACC_ADDITIVE_WEIGHTED, ACC_MULTIPLICATIVE = range(3)
VOXEL_POLAR, VOXEL_UNIFORM = range(2)
OPT_GAUSS_NEWTON, OPT_LM, OPT_SO3_LM = range(3)
SOURCE, TARGET = 0, 1

dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)


class ProjectionParams(C.Structure):
    _fields_ = [("n_scan", C.c_int), ("horizon_scan", C.c_int), ("downsample_rate", C.c_int),
                ("lidar_min_range", C.c_float), ("lidar_max_range", C.c_float),
                ("point_stride_bytes", C.c_int), ("ring_offset_bytes", C.c_int)]


class FeatureParams(C.Structure):
    _fields_ = [("n_scan", C.c_int), ("horizon_scan", C.c_int),
                ("edge_threshold", C.c_float), ("surf_threshold", C.c_float), ("surf_leaf_size", C.c_float)]


class DeskewParams(C.Structure):
    _fields_ = [("time_offset_bytes", C.c_int), ("odom_incre_roll", C.c_float), ("odom_incre_pitch", C.c_float), ("odom_incre_yaw", C.c_float),
                ("scan_period", C.c_float), ("odom_time_diff", C.c_double)]


class MsgHeader(C.Structure):
    _fields_ = [("seq", C.c_uint32), ("stamp_sec", C.c_uint32), ("stamp_nsec", C.c_uint32), ("frame_id", C.c_char_p)]


class CloudInfoScalars(C.Structure):
    _fields_ = [("start_orientation", C.c_float), ("end_orientation", C.c_float), ("orientation_diff", C.c_float), ("initial_guess", C.c_float * 6),
                ("covariance", C.POINTER(C.c_float)), ("n_covariance", C.c_uint32), ("odom_available", C.c_uint8)]


class S2MResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("degenerate", C.c_int32), ("enough_features", C.c_int32)]


REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p)
VOXEL_MERGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)


def declared_symbols() -> list[str]:
    """Every function name declared in include/rolo_gicp.h."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", txt)) - {"rg_reduce_fn"})


_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) the shared library and declare the prototypes used by Python."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    L = C.CDLL(path)
    L.rg_last_error.restype = C.c_char_p
    L.rg_last_error.argtypes = [C.c_void_p]
    L.rg_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.rg_destroy.argtypes = [C.c_void_p]
    L.rg_destroy.restype = None
    for name in ("rg_source_size", "rg_target_size", "rg_get_covariances", "rg_num_correspondences", "rg_voxel_count"):
        getattr(L, name).restype = C.c_size_t
    L.rg_last_kernel_us.restype = C.c_double
    L.rg_batch_last_linearize_us.restype = C.c_double
    L.rg_batch_ctx.restype = C.c_void_p
    L.rg_batch_destroy.restype = None
    L.rg_batch_stream.restype = C.c_void_p
    L.rg_stream.restype = C.c_void_p
    L.rg_kernel_launch_count.restype = C.c_ulonglong
    sz, vp, d, f, i = C.c_size_t, C.c_void_p, C.c_double, C.c_float, C.c_int
    protos = {
        "rg_set_resolution": [vp, d], "rg_set_polar_resolution": [vp, d, d, d], "rg_set_voxel_accumulation_mode": [vp, i],
        "rg_set_neighbor_search_method": [vp, i], "rg_set_num_threads": [vp, i], "rg_set_correspondence_randomness": [vp, i],
        "rg_set_regularization_method": [vp, i], "rg_set_optimizer_type": [vp, i], "rg_set_rotation_epsilon": [vp, d],
        "rg_set_transformation_epsilon": [vp, d], "rg_set_maximum_iterations": [vp, i], "rg_set_initial_lambda_factor": [vp, d],
        "rg_set_debug_print": [vp, i], "rg_set_ct_lambda": [vp, f], "rg_set_profile": [vp, i],
        "rg_set_input_source": [vp, vp, sz, sz], "rg_set_input_target": [vp, vp, sz, sz],
        "rg_set_input_source_device": [vp, vp, sz, vp], "rg_set_input_target_device": [vp, vp, sz, vp],
        "rg_clear_source": [vp], "rg_clear_target": [vp], "rg_swap_source_and_target": [vp],
        "rg_source_size": [vp], "rg_target_size": [vp],
        "rg_set_covariances": [vp, i, vp, sz], "rg_get_covariances": [vp, i, vp],
        "rg_calculate_covariances": [vp, i], "rg_create_voxelmap": [vp],
        "rg_linearize": [vp, dp, dp, dp, dp], "rg_so3_linearize": [vp, dp, dp, dp, dp], "rg_compute_error": [vp, dp, dp],
        "rg_t3_linearize": [vp, dp, dp, dp, d, d, dp, dp, dp], "rg_compute_t_error": [vp, dp, dp, dp, d, d, dp],
        "rg_num_correspondences": [vp],
        "rg_align": [vp, fp, fp], "rg_compute_translation": [vp, dp, dp, dp, d, d, f],
        "rg_transform_source": [vp, fp, vp, sz], "rg_has_converged": [vp], "rg_nr_iterations": [vp],
        "rg_get_final_hessian": [vp, dp], "rg_get_final_pose": [vp, dp],
        "rg_voxel_count": [vp], "rg_get_voxels": [vp, ip, dp, dp, ip], "rg_get_correspondence_keys": [vp, ip],
        "rg_knn": [vp, i, i, ip], "rg_last_kernel_us": [vp],
        "rg_batch_create": [i, i, C.POINTER(vp)], "rg_batch_destroy": [vp], "rg_batch_ctx": [vp, i],
        "rg_batch_align": [vp, fp, fp], "rg_batch_last_linearize_us": [vp], "rg_batch_last_linearize_launches": [vp],
        "rg_set_reduce_hook": [vp, REDUCE_FN, vp, i], "rg_set_source_shard": [vp, sz, sz], "rg_front_end_frame": [vp, vp, sz, vp, vp, i, vp, vp], "rg_cloud_info_serialize": [vp, i, vp, vp, vp, sz, vp], "rg_set_input_from_cloud_info": [vp, i, vp, sz, vp, vp], "rg_s2m_set_map": [vp, vp, sz, vp, sz, sz], "rg_s2m_set_scan": [vp, vp, sz, vp, sz, sz], "rg_s2m_linearize": [vp, vp, vp, vp, vp], "rg_s2m_get_coefficients": [vp, vp, vp, vp, vp], "rg_s2m_optimize": [vp, vp, i, i, i, vp], "rg_sc_make_and_save": [vp, vp, sz, sz, C.c_double], "rg_sc_size": [vp], "rg_sc_clear": [vp], "rg_sc_get": [vp, i, vp, vp, vp], "rg_sc_distance": [vp, i, i, vp, vp], "rg_sc_detect": [vp, vp, vp], "rg_get_fitness_score": [vp, vp, C.c_double, vp], "rg_nearest_target": [vp, vp, vp, vp], "rg_set_deskew": [vp, vp], "rg_shard_export": [vp, vp], "rg_shard_connect": [vp, i, i, vp], "rg_batch_stream": [vp], "rg_stream": [vp], "rg_kernel_launch_count": [],
        "rg_batch_last_stages": [vp, dp], "rg_set_deterministic": [vp, i], "rg_reset_voxelmap": [vp], "rg_set_target_shard": [vp, i, i], "rg_set_voxel_merge_hook": [vp, VOXEL_MERGE_FN, vp], "rg_voxel_import": [vp, vp, vp, i],
    }
    for name, args in protos.items():
        getattr(L, name).argtypes = args
    _lib = L
    return L


class RoloError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rolo_gicp error {code}: {msg}")
        self.code = code


def check(code: int, handle=None):
    """Map C-ABI status codes to the exceptions the reference's C++ layer throws."""
    if code == RG_OK or code == RG_ERR_NOT_CONVERGED:
        return code
    msg = load().rg_last_error(handle)
    msg = msg.decode() if msg else ""
    if code == RG_ERR_INVALID:
        raise ValueError(msg)          # std::invalid_argument
    raise RoloError(code, msg)         # std::runtime_error
