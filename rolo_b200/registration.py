"""Host-side mirror of the reference registration interface over the C-ABI.

`RotVGICP` has the method names, argument meaning and error behaviour of
fast_gicp::RotVGICP<PointT,PointT> (include/rot_gicp/gicp/rot_vgicp.hpp:72-104,
lsq_registration.hpp:48-62, plus the pcl::Registration members lidarOdometry.cpp:460-494 uses),
so tests read like calls against the reference class. All numerics happen in librolo_gicp_b200.so
(CUDA, sm_100a); this file only marshals pointers. Clouds may be numpy arrays (host path, H2D copy
inside the call) or torch CUDA tensors of shape (N,4) float32 (device-resident path).
"""
from __future__ import annotations

import ctypes as C
from enum import IntEnum

import numpy as np

from . import capi


class RegularizationMethod(IntEnum):   # gicp_settings.hpp:6
    NONE = 0; MIN_EIG = 1; NORMALIZED_MIN_EIG = 2; PLANE = 3; FROBENIUS = 4; PLANE_S = 5


class NeighborSearchMethod(IntEnum):   # gicp_settings.hpp:8
    DIRECT27 = 0; DIRECT7 = 1; DIRECT1 = 2; DIRECT_RADIUS = 3


class VoxelAccumulationMode(IntEnum):  # gicp_settings.hpp:10
    ADDITIVE = 0; ADDITIVE_WEIGHTED = 1; MULTIPLICATIVE = 2


class LSQ_OPTIMIZER_TYPE(IntEnum):     # lsq_registration.hpp:13
    GaussNewton = 0; LevenbergMarquardt = 1; SO3_LevenbergMarquardt = 2


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(capi.dp)


class RotVGICP:
    """Drop-in counterpart of fast_gicp::RotVGICP (rot_vgicp.hpp:24)."""

    def __init__(self, device: int = -1, _handle=None, _owner=None):
        self._L = capi.load()
        self._owner = _owner
        if _handle is not None:
            self._h = C.c_void_p(_handle)
        else:
            h = C.c_void_p()
            capi.check(self._L.rg_create(device, C.byref(h)))
            self._h = h
        self._keep = {}
        self._final = np.eye(4, dtype=np.float32)
        self._rc = 0

    def __del__(self):
        try:
            if self._owner is None and self._h:
                self._L.rg_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _ck(self, code):
        return capi.check(code, self._h)

    # ---- setters ------------------------------------------------------------------------
    def setResolution(self, resolution: float): self._ck(self._L.rg_set_resolution(self._h, resolution))
    def setPolarResolution(self, theta_res, phi_res, r_res): self._ck(self._L.rg_set_polar_resolution(self._h, theta_res, phi_res, r_res))
    def setVoxelAccumulationMode(self, mode): self._ck(self._L.rg_set_voxel_accumulation_mode(self._h, int(mode)))
    def setNeighborSearchMethod(self, method): self._ck(self._L.rg_set_neighbor_search_method(self._h, int(method)))
    def setNumThreads(self, n): self._ck(self._L.rg_set_num_threads(self._h, int(n)))
    def setCorrespondenceRandomness(self, k): self._ck(self._L.rg_set_correspondence_randomness(self._h, int(k)))
    def setRegularizationMethod(self, method): self._ck(self._L.rg_set_regularization_method(self._h, int(method)))
    def setOptimizerType(self, t): self._ck(self._L.rg_set_optimizer_type(self._h, int(t)))
    def setRotationEpsilon(self, eps): self._ck(self._L.rg_set_rotation_epsilon(self._h, eps))
    def setTransformationEpsilon(self, eps): self._ck(self._L.rg_set_transformation_epsilon(self._h, eps))
    def setMaximumIterations(self, n): self._ck(self._L.rg_set_maximum_iterations(self._h, int(n)))
    def setInitialLambdaFactor(self, f): self._ck(self._L.rg_set_initial_lambda_factor(self._h, f))
    def setDebugPrint(self, on): self._ck(self._L.rg_set_debug_print(self._h, int(bool(on))))
    def setProfile(self, on): self._ck(self._L.rg_set_profile(self._h, int(bool(on))))

    # ---- clouds -----------------------------------------------------------------------
    def _set_cloud(self, which, cloud):
        fn_host = self._L.rg_set_input_source if which == capi.SOURCE else self._L.rg_set_input_target
        fn_dev = self._L.rg_set_input_source_device if which == capi.SOURCE else self._L.rg_set_input_target_device
        if _is_torch_cuda(cloud):
            import torch
            assert cloud.dtype == torch.float32 and cloud.dim() == 2 and cloud.shape[1] == 4 and cloud.is_contiguous()
            stream = torch.cuda.current_stream(cloud.device).cuda_stream
            self._ck(fn_dev(self._h, C.c_void_p(cloud.data_ptr()), cloud.shape[0], C.c_void_p(stream)))
            self._keep[which] = cloud
            return
        if hasattr(cloud, "numpy") and not isinstance(cloud, np.ndarray):
            cloud = cloud.numpy()   # CPU torch tensor (e.g. pinned): zero-copy view
        a = np.asarray(cloud)
        if a.dtype != np.float32 or a.ndim != 2 or a.shape[1] < 3 or not a.flags.c_contiguous:
            a = np.ascontiguousarray(a, dtype=np.float32)
        self._ck(fn_host(self._h, C.c_void_p(a.ctypes.data), a.shape[0], a.strides[0]))
        self._keep[which] = a

    def setInputSource(self, cloud): self._set_cloud(capi.SOURCE, cloud)
    def setInputTarget(self, cloud): self._set_cloud(capi.TARGET, cloud)
    def clearSource(self): self._ck(self._L.rg_clear_source(self._h))
    def clearTarget(self): self._ck(self._L.rg_clear_target(self._h))
    def swapSourceAndTarget(self): self._ck(self._L.rg_swap_source_and_target(self._h))

    def _set_covs(self, which, covs):
        a = np.ascontiguousarray(covs, dtype=np.float64).reshape(-1, 16)
        self._ck(self._L.rg_set_covariances(self._h, which, C.c_void_p(a.ctypes.data), a.shape[0]))

    def setSourceCovariances(self, covs): self._set_covs(capi.SOURCE, covs)
    def setTargetCovariances(self, covs): self._set_covs(capi.TARGET, covs)

    def _get_covs(self, which):
        n = self._L.rg_get_covariances(self._h, which, None)
        out = np.zeros((n, 4, 4), dtype=np.float64)
        if n:
            got = self._L.rg_get_covariances(self._h, which, C.c_void_p(out.ctypes.data))
            if got != n:
                raise capi.RoloError(capi.RG_ERR_CUDA, "rg_get_covariances failed")
        return out

    def getSourceCovariances(self): return self._get_covs(capi.SOURCE)
    def getTargetCovariances(self): return self._get_covs(capi.TARGET)

    # ---- stages (protected virtuals of the reference; public here for the parity tests) --------
    def calculate_covariances(self, which): self._ck(self._L.rg_calculate_covariances(self._h, which))
    def create_voxelmap(self): self._ck(self._L.rg_create_voxelmap(self._h))

    @staticmethod
    def _pose(T):
        return np.ascontiguousarray(T, dtype=np.float64).reshape(16)

    def linearize(self, trans, want_H=True):
        T = self._pose(trans); cost = C.c_double()
        if not want_H:
            self._ck(self._L.rg_linearize(self._h, _dptr(T), None, None, C.byref(cost)))
            return cost.value, None, None
        H = np.empty((6, 6)); b = np.empty(6)
        self._ck(self._L.rg_linearize(self._h, _dptr(T), _dptr(H), _dptr(b), C.byref(cost)))
        return cost.value, H, b

    def so3_linearize(self, trans):
        T = self._pose(trans); cost = C.c_double(); H = np.empty((3, 3)); b = np.empty(3)
        self._ck(self._L.rg_so3_linearize(self._h, _dptr(T), _dptr(H), _dptr(b), C.byref(cost)))
        return cost.value, H, b

    def compute_error(self, trans):
        T = self._pose(trans); cost = C.c_double()
        self._ck(self._L.rg_compute_error(self._h, _dptr(T), C.byref(cost)))
        return cost.value

    def t3_linearize(self, trans, init_guess, last_t0, interval_tn, interval_tn_1):
        t, g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (trans, init_guess, last_t0))
        cost = C.c_double(); H = np.empty((6, 6)); b = np.empty(6)
        self._ck(self._L.rg_t3_linearize(self._h, _dptr(t), _dptr(g), _dptr(l), interval_tn, interval_tn_1, _dptr(H), _dptr(b), C.byref(cost)))
        return cost.value, H, b

    def compute_t_error(self, trans, init_guess, last_t0, interval_tn, interval_tn_1):
        t, g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (trans, init_guess, last_t0))
        cost = C.c_double()
        self._ck(self._L.rg_compute_t_error(self._h, _dptr(t), _dptr(g), _dptr(l), interval_tn, interval_tn_1, C.byref(cost)))
        return cost.value

    def evaluateCost(self, relative_pose, want_H=False):   # lsq_registration_impl.hpp:50-52
        return self.linearize(np.asarray(relative_pose, dtype=np.float32).astype(np.float64), want_H)

    def setCtLambda(self, lam): self._ck(self._L.rg_set_ct_lambda(self._h, lam))
    def numCorrespondences(self): return self._L.rg_num_correspondences(self._h)

    # ---- drivers ------------------------------------------------------------------------------
    def align(self, guess=None, output=None):
        """pcl::Registration::align(output[, guess]); returns the final 4x4 float transformation.
        `output` (optional numpy (N,>=3) float32) receives the transformed source like the reference's
        output cloud; it must not alias the source or target buffers (rot_vgicp_impl.hpp:150-152)."""
        if output is not None:
            for w in (capi.SOURCE, capi.TARGET):
                k = self._keep.get(w)
                if isinstance(k, np.ndarray) and isinstance(output, np.ndarray) and np.shares_memory(k, output):
                    raise ValueError("RotVGICP: destination cloud cannot be identical to source or target")
        g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float32).reshape(16)
        out = np.empty(16, dtype=np.float32)
        self._rc = self._ck(self._L.rg_align(self._h, None if g is None else g.ctypes.data_as(capi.fp), out.ctypes.data_as(capi.fp)))
        self._final = out.reshape(4, 4)
        if output is not None:
            self._ck(self._L.rg_transform_source(self._h, out.ctypes.data_as(capi.fp), C.c_void_p(output.ctypes.data), output.strides[0]))
        return self._final

    def computeTranslation(self, trans, init_guess, last_t0, interval_tn, interval_tn_1, ct_lambda):
        t = np.array(trans, dtype=np.float64)
        g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (init_guess, last_t0))
        self._rc = self._ck(self._L.rg_compute_translation(self._h, _dptr(t), _dptr(g), _dptr(l), interval_tn, interval_tn_1, ct_lambda))
        return t

    def getFinalTransformation(self): return self._final
    def hasConverged(self): return bool(self._L.rg_has_converged(self._h))
    def nrIterations(self): return self._L.rg_nr_iterations(self._h)
    def lmFailed(self): return self._rc == capi.RG_ERR_NOT_CONVERGED

    def getFinalHessian(self):
        H = np.empty((6, 6)); self._ck(self._L.rg_get_final_hessian(self._h, _dptr(H))); return H

    def getFinalPoseDouble(self):
        T = np.empty((4, 4)); self._ck(self._L.rg_get_final_pose(self._h, _dptr(T))); return T

    # ---- introspection for tests ---------------------------------------------------------------
    def getVoxels(self):
        v = self._L.rg_voxel_count(self._h)
        keys = np.empty((v, 3), dtype=np.int32); mean = np.empty((v, 3)); cov6 = np.empty((v, 6)); npts = np.empty(v, dtype=np.int32)
        self._ck(self._L.rg_get_voxels(self._h, keys.ctypes.data_as(capi.ip), _dptr(mean), _dptr(cov6), npts.ctypes.data_as(capi.ip)))
        return keys, mean, cov6, npts

    def getCorrespondenceKeys(self, noff=1):
        n = self._L.rg_source_size(self._h)
        keys = np.empty((n * noff, 3), dtype=np.int32)
        self._ck(self._L.rg_get_correspondence_keys(self._h, keys.ctypes.data_as(capi.ip)))
        return keys

    def knn(self, which, k):
        n = self._L.rg_source_size(self._h) if which == capi.SOURCE else self._L.rg_target_size(self._h)
        out = np.empty((n, k), dtype=np.int32)
        self._ck(self._L.rg_knn(self._h, which, k, out.ctypes.data_as(capi.ip)))
        return out

    def lastKernelUs(self): return self._L.rg_last_kernel_us(self._h)
    def stream(self): return self._L.rg_stream(self._h)

    def frontEndFrame(self, raw, which, n_scan=128, horizon_scan=2048, downsample_rate=1, lidar_min_range=2.0, lidar_max_range=1000.0,
                      edge_threshold=0.8, surf_threshold=0.1, surf_leaf_size=0.4):
        """Raw driver frame -> range image -> features -> `corner + surface` installed as this object's target/source cloud,
        all on the device (imageProjection.cpp:399-505, featureExtraction.cpp:87-266, lidarOdometry.cpp:521-524)."""
        raw = np.ascontiguousarray(raw)
        pp = capi.ProjectionParams(n_scan, horizon_scan, downsample_rate, lidar_min_range, lidar_max_range, raw.dtype.itemsize, 20)
        fp = capi.FeatureParams(n_scan, horizon_scan, edge_threshold, surf_threshold, surf_leaf_size)
        nc = C.c_int32(); ns = C.c_int32()
        self._ck(self._L.rg_front_end_frame(self._h, C.c_void_p(raw.ctypes.data), raw.shape[0], C.byref(pp), C.byref(fp), which, C.byref(nc), C.byref(ns)))
        self._keep[which] = None
        return nc.value, ns.value

    def cloudInfoSerialize(self, stage: int, seq: int = 0, stamp_sec: int = 0, stamp_nsec: int = 0, frame_id: str = "", scalars9=None,
                           covariance=None, odom_available: bool = False) -> bytes:
        """ROS1 wire image of rolo/CloudInfoStamp for the frame last run through frontEndFrame: stage 0 as imageProjection
        publishes it (imageProjection.cpp:507-512), stage 1 as featureExtraction does (featureExtraction.cpp:268-287)."""
        hdr = capi.MsgHeader(seq, stamp_sec, stamp_nsec, frame_id.encode())
        sc = capi.CloudInfoScalars()
        s9 = np.zeros(9, np.float32) if scalars9 is None else np.asarray(scalars9, np.float32)
        sc.start_orientation, sc.end_orientation, sc.orientation_diff = float(s9[0]), float(s9[1]), float(s9[2])
        for k in range(6): sc.initial_guess[k] = float(s9[3 + k])
        cov = np.zeros(0, np.float32) if covariance is None else np.ascontiguousarray(covariance, np.float32)
        sc.covariance = cov.ctypes.data_as(C.POINTER(C.c_float)) if cov.size else None
        sc.n_covariance = cov.size; sc.odom_available = 1 if odom_available else 0
        n = C.c_size_t()
        self._ck(self._L.rg_cloud_info_serialize(self._h, stage, C.byref(hdr), C.byref(sc), None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint8)
        self._ck(self._L.rg_cloud_info_serialize(self._h, stage, C.byref(hdr), C.byref(sc), C.c_void_p(out.ctypes.data), n.value, C.byref(n)))
        return out.tobytes()

    def setInputFromCloudInfo(self, which: int, msg: bytes):
        """Consumer edge (lidarOdometry.cpp:521-524): corner + surface of a serialised CloudInfoStamp become the source/target."""
        a = np.frombuffer(msg, dtype=np.uint8)
        nc = C.c_int32(); ns = C.c_int32()
        self._ck(self._L.rg_set_input_from_cloud_info(self._h, which, C.c_void_p(a.ctypes.data), a.size, C.byref(nc), C.byref(ns)))
        self._keep[which] = None
        return nc.value, ns.value

    def getFitnessScore(self, T=None, max_range: float = float(np.finfo(np.float64).max)) -> float:
        """pcl::Registration::getFitnessScore(max_range) at transform T (default: the final transformation of the last align)."""
        T = self._final if T is None else T
        t = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
        s = C.c_double()
        self._ck(self._L.rg_get_fitness_score(self._h, t.ctypes.data_as(capi.fp), max_range, C.byref(s)))
        return s.value

    def nearestTarget(self, T, n_source: int):
        t = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
        idx = np.empty(n_source, np.int32); d2 = np.empty(n_source, np.float32)
        self._ck(self._L.rg_nearest_target(self._h, t.ctypes.data_as(capi.fp), C.c_void_p(idx.ctypes.data), C.c_void_p(d2.ctypes.data)))
        return idx, d2

    def setSourceShard(self, lo: int, hi: int): self._ck(self._L.rg_set_source_shard(self._h, lo, hi))

    def shardExport(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self._L.rg_shard_export(self._h, buf))
        return buf.raw

    def shardConnect(self, rank: int, world: int, handles: list):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._ck(self._L.rg_shard_connect(self._h, rank, world, blob))

    def setDeterministic(self, on: bool = True):
        """Sort + fixed-order voxel sums (bit-reproducible voxel map) instead of fp64 atomics."""
        self._ck(self._L.rg_set_deterministic(self._h, int(bool(on))))

    def setTargetShard(self, rank: int, world: int):
        """Target-sharded setup (config 4): this rank estimates covariances / accumulates voxels for its slice of the target only."""
        self._ck(self._L.rg_set_target_shard(self._h, int(rank), int(world)))

    def setVoxelMergeHook(self, fn):
        """fn(keys_ptr:int, sums_ptr:int, n_local:int, stream:int) -> int; must call voxelImport() with the list gathered over all ranks."""
        if fn is None:
            self._merge_hook = None
            self._ck(self._L.rg_set_voxel_merge_hook(self._h, capi.VOXEL_MERGE_FN(0), None))
            return

        def tramp(user, keys, sums, n_local, stream):
            try:
                return int(fn(keys or 0, sums or 0, int(n_local), stream or 0) or 0)
            except Exception as e:   # never let an exception cross the C boundary
                print("voxel merge hook failed:", repr(e))
                return -1
        self._merge_hook = capi.VOXEL_MERGE_FN(tramp)
        self._ck(self._L.rg_set_voxel_merge_hook(self._h, self._merge_hook, None))

    def voxelImport(self, keys_ptr: int, sums_ptr: int, n_total: int):
        self._ck(self._L.rg_voxel_import(self._h, C.c_void_p(keys_ptr), C.c_void_p(sums_ptr), int(n_total)))

    def setReduceHook(self, fn, on_device: bool):
        """Sharded registration (config 4): fn(buf_ptr:int, count:int, on_device:bool, stream:int) -> int."""
        if fn is None:
            self._hook = None
            self._ck(self._L.rg_set_reduce_hook(self._h, capi.REDUCE_FN(0), None, 0))
            return

        def tramp(user, buf, count, on_dev, stream):
            try:
                return int(fn(buf, count, bool(on_dev), stream or 0) or 0)
            except Exception as e:   # never let an exception cross the C boundary
                print("reduce hook failed:", e)
                return -1
        self._hook = capi.REDUCE_FN(tramp)
        self._ck(self._L.rg_set_reduce_hook(self._h, self._hook, None, int(on_device)))


class RotVGICPBatch:
    """A lock-step batch of independent scan pairs on one GPU (BASELINE config 3)."""

    def __init__(self, n_pairs: int, device: int = -1):
        self._L = capi.load()
        h = C.c_void_p()
        capi.check(self._L.rg_batch_create(device, n_pairs, C.byref(h)))
        self._h = h
        self.n_pairs = n_pairs
        self.pairs = [RotVGICP(_handle=self._L.rg_batch_ctx(self._h, i), _owner=self) for i in range(n_pairs)]

    def __del__(self):
        try:
            if self._h:
                self._L.rg_batch_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def __getitem__(self, i): return self.pairs[i]

    def align(self, guesses=None):
        g = None if guesses is None else np.ascontiguousarray(guesses, dtype=np.float32).reshape(self.n_pairs * 16)
        out = np.empty((self.n_pairs, 4, 4), dtype=np.float32)
        rc = self._L.rg_batch_align(self._h, None if g is None else g.ctypes.data_as(capi.fp), out.ctypes.data_as(capi.fp))
        capi.check(rc)
        for i, p in enumerate(self.pairs):
            p._final = out[i]
        return out

    def stream(self): return self._L.rg_batch_stream(self._h)

    def lastStages(self) -> dict:
        """Stage times (ms) of the last align(): CUDA events on the batch stream."""
        out = np.zeros(8)
        capi.check(self._L.rg_batch_last_stages(self._h, out.ctypes.data_as(capi.dp)))
        return {"covariances_ms": out[0], "knn_kernels_ms": out[1], "voxel_maps_ms": out[2], "iterations_ms": out[3], "linearize_kernels_ms": out[4],
                "covariance_points": int(out[5]), "linearize_launches": int(out[6])}
    def lastLinearizeUs(self): return self._L.rg_batch_last_linearize_us(self._h)
    def lastLinearizeLaunches(self): return self._L.rg_batch_last_linearize_launches(self._h)
