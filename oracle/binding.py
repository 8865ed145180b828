"""ctypes binding of the CPU ORACLE (oracle/rolo_oracle.h). Test infrastructure only.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package rolo_b200/ must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librolo_oracle.so")

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS, REG_PLANE_S = range(6)
DIRECT27, DIRECT7, DIRECT1, DIRECT_RADIUS = range(4)
VOXEL_POLAR, VOXEL_UNIFORM = range(2)
OPT_GAUSS_NEWTON, OPT_LM, OPT_SO3_LM = range(3)

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("rolo_oracle.cpp", "rolo_oracle.h", "frontend_oracle.cpp", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.ro_create.restype = C.c_void_p
        for name in ("ro_linearize", "ro_so3_linearize", "ro_compute_error", "ro_t3_linearize", "ro_compute_t_error"):
            getattr(L, name).restype = C.c_double
        for name in ("ro_get_covariances", "ro_voxel_count", "ro_num_correspondences", "ro_trace_size"):
            getattr(L, name).restype = C.c_size_t
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


class Oracle:
    """Thin object wrapper; method names follow the reference class (rot_vgicp.hpp:72-104)."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.ro_create())
        self._keep = []

    def __del__(self):
        try:
            self.L.ro_destroy(self.h)
        except Exception:
            pass

    # --- configuration -------------------------------------------------------------
    def setNumThreads(self, n): self.L.ro_set_num_threads(self.h, C.c_int(n))
    def getNumThreads(self): return self.L.ro_get_num_threads(self.h)
    def setResolution(self, r): self.L.ro_set_resolution(self.h, C.c_double(r))
    def setPolarResolution(self, th, ph, r): self.L.ro_set_polar_resolution(self.h, C.c_double(th), C.c_double(ph), C.c_double(r))
    def setNeighborSearchMethod(self, m): self.L.ro_set_neighbor_search_method(self.h, C.c_int(m))
    def setRegularizationMethod(self, m): self.L.ro_set_regularization_method(self.h, C.c_int(m))
    def setCorrespondenceRandomness(self, k): self.L.ro_set_correspondence_randomness(self.h, C.c_int(k))
    def setOptimizerType(self, t): self.L.ro_set_optimizer_type(self.h, C.c_int(t))
    def setMaximumIterations(self, n): self.L.ro_set_max_iterations(self.h, C.c_int(n))
    def setRotationEpsilon(self, e): self.L.ro_set_rotation_epsilon(self.h, C.c_double(e))
    def setTransformationEpsilon(self, e): self.L.ro_set_transformation_epsilon(self.h, C.c_double(e))
    def setInitialLambdaFactor(self, f): self.L.ro_set_initial_lambda_factor(self.h, C.c_double(f))

    # --- clouds ----------------------------------------------------------------------
    def _cloud(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        assert pts.ndim == 2 and pts.shape[1] >= 3
        return pts

    def setInputSource(self, pts):
        pts = self._cloud(pts)
        self.L.ro_set_input_source(self.h, pts.ctypes.data_as(_fp), C.c_size_t(pts.shape[0]), C.c_size_t(pts.shape[1]))
        self.n_src = pts.shape[0]

    def setInputTarget(self, pts):
        pts = self._cloud(pts)
        self.L.ro_set_input_target(self.h, pts.ctypes.data_as(_fp), C.c_size_t(pts.shape[0]), C.c_size_t(pts.shape[1]))
        self.n_tgt = pts.shape[0]

    def setSourceCovariances(self, covs): self._set_covs(0, covs)
    def setTargetCovariances(self, covs): self._set_covs(1, covs)

    def _set_covs(self, which, covs):
        covs = np.ascontiguousarray(covs, dtype=np.float64).reshape(-1, 16)
        self.L.ro_set_covariances(self.h, C.c_int(which), _d(covs), C.c_size_t(covs.shape[0]))

    def calculateCovariances(self, which):
        return self.L.ro_calculate_covariances(self.h, C.c_int(which))

    def getCovariances(self, which):
        n = self.L.ro_get_covariances(self.h, C.c_int(which), None)
        out = np.empty((n, 4, 4), dtype=np.float64)
        self.L.ro_get_covariances(self.h, C.c_int(which), _d(out))
        return out

    def getSourceCovariances(self): return self.getCovariances(0)
    def getTargetCovariances(self): return self.getCovariances(1)

    def knn(self, which, k):
        n = self.n_src if which == 0 else self.n_tgt
        out = np.empty((n, k), dtype=np.int32)
        self.L.ro_knn(self.h, C.c_int(which), C.c_int(k), out.ctypes.data_as(_ip))
        return out

    # --- voxel map ---------------------------------------------------------------------
    def createVoxelmap(self): self.L.ro_create_voxelmap(self.h)
    def resetVoxelmap(self): self.L.ro_reset_voxelmap(self.h)

    def getVoxels(self):
        v = self.L.ro_voxel_count(self.h)
        keys = np.empty((v, 3), dtype=np.int32)
        mean = np.empty((v, 4), dtype=np.float64)
        cov = np.empty((v, 4, 4), dtype=np.float64)
        npts = np.empty((v,), dtype=np.int32)
        self.L.ro_get_voxels(self.h, keys.ctypes.data_as(_ip), _d(mean), _d(cov), npts.ctypes.data_as(_ip))
        return keys, mean, cov, npts

    # --- hot-path evaluations ------------------------------------------------------------
    @staticmethod
    def _T(T):
        return np.ascontiguousarray(T, dtype=np.float64).reshape(16)

    def linearize(self, T, want_H=True):
        T = self._T(T)
        if not want_H:
            return self.L.ro_linearize(self.h, _d(T), None, None), None, None
        H = np.empty((6, 6)); b = np.empty(6)
        c = self.L.ro_linearize(self.h, _d(T), _d(H), _d(b))
        return c, H, b

    def so3_linearize(self, T):
        T = self._T(T)
        H = np.empty((3, 3)); b = np.empty(3)
        c = self.L.ro_so3_linearize(self.h, _d(T), _d(H), _d(b))
        return c, H, b

    def compute_error(self, T):
        T = self._T(T)
        return self.L.ro_compute_error(self.h, _d(T))

    def t3_linearize(self, t, guess, last_t0, dtn, dtn1):
        t, g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (t, guess, last_t0))
        H = np.empty((6, 6)); b = np.empty(6)
        c = self.L.ro_t3_linearize(self.h, _d(t), _d(g), _d(l), C.c_double(dtn), C.c_double(dtn1), _d(H), _d(b))
        return c, H, b

    def compute_t_error(self, t, guess, last_t0, dtn, dtn1):
        t, g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (t, guess, last_t0))
        return self.L.ro_compute_t_error(self.h, _d(t), _d(g), _d(l), C.c_double(dtn), C.c_double(dtn1))

    def setCtLambda(self, lam): self.L.ro_set_ct_lambda(self.h, C.c_float(lam))

    def getCorrespondences(self):
        m = self.L.ro_num_correspondences(self.h)
        idx = np.empty((m,), dtype=np.int32); key = np.empty((m, 3), dtype=np.int32)
        self.L.ro_get_correspondences(self.h, idx.ctypes.data_as(_ip), key.ctypes.data_as(_ip))
        return idx, key

    # --- drivers ---------------------------------------------------------------------------
    def align(self, guess=None):
        g = np.eye(4, dtype=np.float32) if guess is None else np.ascontiguousarray(guess, dtype=np.float32)
        out = np.empty((4, 4), dtype=np.float32)
        self.rc = self.L.ro_align(self.h, g.ctypes.data_as(_fp), out.ctypes.data_as(_fp))
        self.final_transformation = out
        return out

    def getFinalTransformation(self): return self.final_transformation

    def getFinalPoseDouble(self):
        T = np.empty((4, 4)); self.L.ro_get_final_pose_d(self.h, _d(T)); return T

    def getFitnessScore(self, T, max_range=float(np.finfo(np.float64).max), n_source=None):
        """pcl::Registration::getFitnessScore(max_range); with n_source also returns (idx, d2) per source point."""
        self.L.ro_get_fitness_score.restype = C.c_double
        t = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
        if n_source is None:
            return self.L.ro_get_fitness_score(self.h, C.c_void_p(t.ctypes.data), C.c_double(max_range), None, None)
        idx = np.empty(n_source, np.int32); d2 = np.empty(n_source, np.float32)
        sc = self.L.ro_get_fitness_score(self.h, C.c_void_p(t.ctypes.data), C.c_double(max_range), C.c_void_p(idx.ctypes.data), C.c_void_p(d2.ctypes.data))
        return sc, idx, d2

    def hasConverged(self): return bool(self.L.ro_has_converged(self.h))
    def nrIterations(self): return self.L.ro_nr_iterations(self.h)

    def getFinalHessian(self):
        H = np.empty((6, 6)); self.L.ro_get_final_hessian(self.h, _d(H)); return H

    def computeTranslation(self, trans, init_guess, last_t0, dtn, dtn1, ct_lambda):
        t = np.array(trans, dtype=np.float64)
        g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (init_guess, last_t0))
        self.rc = self.L.ro_compute_translation(self.h, _d(t), _d(g), _d(l), C.c_double(dtn), C.c_double(dtn1), C.c_float(ct_lambda))
        return t

    def getTrace(self):
        n = self.L.ro_trace_size(self.h)
        out = np.empty((n, 64), dtype=np.float64)
        if n:
            self.L.ro_get_trace(self.h, _d(out))
        return out


def so3_exp(w):
    w = np.ascontiguousarray(w, dtype=np.float64); q = np.empty(4)
    lib().ro_so3_exp(_d(w), _d(q)); return q


def se3_exp(a):
    a = np.ascontiguousarray(a, dtype=np.float64); T = np.empty((4, 4))
    lib().ro_se3_exp(_d(a), _d(T)); return T


def svd3(A):
    A = np.ascontiguousarray(A, dtype=np.float64); U = np.empty((3, 3)); S = np.empty(3); V = np.empty((3, 3))
    lib().ro_svd3(_d(A), _d(U), _d(S), _d(V)); return U, S, V


def ldlt_solve(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b)
    lib().ro_ldlt_solve(_d(A), _d(b), C.c_int(b.shape[0]), _d(x)); return x


# ---- upstream range-image passes (oracle/frontend_oracle.cpp) ---------------------------------------
def image_projection(raw: np.ndarray, n_scan, horizon, downsample_rate=1, min_range=2.0, max_range=1000.0, ring_offset=20, deskew=None):
    """deskew = None (deskewEnabled false / no odometry) or dict(time_offset, incre_rpy(3), scan_period, odom_time_diff)."""
    L = lib()
    raw = np.ascontiguousarray(raw)
    cells = n_scan * horizon
    rim = np.empty((n_scan, horizon), dtype=np.float32)
    ext = np.empty((cells, 4), dtype=np.float32); col = np.empty(cells, dtype=np.int32); rng = np.empty(cells, dtype=np.float32)
    st = np.empty(n_scan, dtype=np.int32); en = np.empty(n_scan, dtype=np.int32)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    if deskew is None:
        n = L.ro_image_projection(vp(raw), C.c_size_t(raw.shape[0]), C.c_int(raw.dtype.itemsize), C.c_int(ring_offset), C.c_int(n_scan), C.c_int(horizon),
                                  C.c_int(downsample_rate), C.c_float(min_range), C.c_float(max_range), vp(rim), vp(ext), vp(col), vp(rng), vp(st), vp(en))
    else:
        inc = np.ascontiguousarray(deskew["incre_rpy"], dtype=np.float32)
        n = L.ro_image_projection_deskew(vp(raw), C.c_size_t(raw.shape[0]), C.c_int(raw.dtype.itemsize), C.c_int(ring_offset), C.c_int(n_scan), C.c_int(horizon),
                                         C.c_int(downsample_rate), C.c_float(min_range), C.c_float(max_range), C.c_int(1), C.c_int(deskew["time_offset"]),
                                         vp(inc), C.c_float(deskew["scan_period"]), C.c_double(deskew["odom_time_diff"]),
                                         vp(rim), vp(ext), vp(col), vp(rng), vp(st), vp(en))
    return {"range_image": rim, "extracted": ext[:n].copy(), "point_col_ind": col[:n].copy(), "point_range": rng[:n].copy(),
            "start_ring_index": st, "end_ring_index": en, "n": n}


def feature_extraction(proj: dict, n_scan, horizon, edge_threshold=0.8, surf_threshold=0.1, surf_leaf_size=0.4):
    L = lib()
    n = proj["n"]
    ext = np.ascontiguousarray(proj["extracted"], dtype=np.float32); col = np.ascontiguousarray(proj["point_col_ind"], dtype=np.int32)
    rng = np.ascontiguousarray(proj["point_range"], dtype=np.float32)
    st = np.ascontiguousarray(proj["start_ring_index"], dtype=np.int32); en = np.ascontiguousarray(proj["end_ring_index"], dtype=np.int32)
    curv = np.empty(n, dtype=np.float32); picked = np.empty(n, dtype=np.int32); label = np.empty(n, dtype=np.int32)
    corner = np.empty((n_scan * 120, 4), dtype=np.float32); surf = np.empty((n_scan * horizon, 4), dtype=np.float32)
    nc = C.c_int(); ns = C.c_int()
    vp = lambda a: C.c_void_p(a.ctypes.data)
    L.ro_feature_extraction(vp(ext), vp(col), vp(rng), vp(st), vp(en), C.c_int(n), C.c_int(n_scan), C.c_int(horizon), C.c_float(edge_threshold),
                            C.c_float(surf_threshold), C.c_float(surf_leaf_size), vp(curv), vp(picked), vp(label), vp(corner), C.byref(nc), vp(surf), C.byref(ns))
    return {"curvature": curv, "neighbor_picked": picked, "label": label, "corner": corner[:nc.value].copy(), "surface": surf[:ns.value].copy()}


def cloud_info_serialize(stage: int, seq: int, sec: int, nsec: int, frame: str, proj: dict, feat: dict | None = None,
                         scalars9=None, covariance=None, odom_available: bool = False) -> bytes:
    """ROS1 wire image of rolo/CloudInfoStamp as imageProjection (stage 0) / featureExtraction (stage 1) publish it
    (cloud_info_oracle.cpp). `proj` / `feat` are image_projection() / feature_extraction() results."""
    L = lib()
    L.ro_cloud_info_serialize.restype = C.c_size_t
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)
    st, en, col, rng, ext = i32(proj["start_ring_index"]), i32(proj["end_ring_index"]), i32(proj["point_col_ind"]), f32(proj["point_range"]), f32(proj["extracted"])
    corner = f32(feat["corner"]) if feat is not None else np.zeros((0, 4), np.float32)
    surf = f32(feat["surface"]) if feat is not None else np.zeros((0, 4), np.float32)
    sc = f32(scalars9) if scalars9 is not None else np.zeros(9, np.float32)
    cov = f32(covariance) if covariance is not None else np.zeros(0, np.float32)
    args = lambda out, cap: (C.c_int(stage), C.c_uint32(seq), C.c_uint32(sec), C.c_uint32(nsec), C.c_char_p(frame.encode()), vp(st), vp(en), C.c_int(st.shape[0]),
                             vp(col), vp(rng), vp(ext), C.c_int(ext.shape[0]), vp(corner), C.c_int(corner.shape[0]), vp(surf), C.c_int(surf.shape[0]),
                             vp(sc), vp(cov), C.c_int(cov.shape[0]), C.c_int(1 if odom_available else 0), out, C.c_size_t(cap))
    size = L.ro_cloud_info_serialize(*args(C.c_void_p(0), 0))
    buf = np.empty(size, dtype=np.uint8)
    L.ro_cloud_info_serialize(*args(C.c_void_p(buf.ctypes.data), size))
    return buf.tobytes()


def cloud_info_features(msg: bytes) -> tuple[np.ndarray, int, int]:
    """lidarOdometry.cpp:521-524: featureLast = fromROSMsg(extracted_corner) + fromROSMsg(extracted_surface), (n, 4) xyzi."""
    L = lib()
    a = np.frombuffer(msg, dtype=np.uint8)
    nc, ns = C.c_int(), C.c_int()
    if L.ro_cloud_info_features(C.c_void_p(a.ctypes.data), C.c_size_t(a.size), C.c_void_p(0), C.byref(nc), C.byref(ns)) != 0:
        raise ValueError("malformed CloudInfoStamp")
    out = np.empty((nc.value + ns.value, 4), dtype=np.float32)
    L.ro_cloud_info_features(C.c_void_p(a.ctypes.data), C.c_size_t(a.size), C.c_void_p(out.ctypes.data), C.byref(nc), C.byref(ns))
    return out, nc.value, ns.value


class Scan2Map:
    """backMapping's scan-to-map optimisation (scan2map_oracle.cpp; backMapping.cpp:681-1054). Clouds are (n, >=3) float32."""

    def __init__(self, threads: int = 0):
        self.L = lib()
        self.L.ro_s2m_create.restype = C.c_void_p
        self.h = C.c_void_p(self.L.ro_s2m_create())
        self.L.ro_s2m_set_num_threads(self.h, threads)
        self.n_corner = self.n_surf = 0

    def __del__(self):
        try:
            self.L.ro_s2m_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def _c(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        return a, C.c_void_p(a.ctypes.data), C.c_size_t(a.shape[0]), a.shape[1]

    def setMap(self, corner, surf):
        c, cp, cn, cs = self._c(corner); s, sp, sn, ss = self._c(surf)
        assert cs == ss
        self.L.ro_s2m_set_map(self.h, cp, cn, sp, sn, C.c_size_t(cs))

    def setScan(self, corner, surf):
        c, cp, cn, cs = self._c(corner); s, sp, sn, ss = self._c(surf)
        assert cs == ss
        self.n_corner, self.n_surf = c.shape[0], s.shape[0]
        self.L.ro_s2m_set_scan(self.h, cp, cn, sp, sn, C.c_size_t(cs))

    def linearize(self, tf6):
        tf = np.ascontiguousarray(tf6, dtype=np.float32)
        AtA = np.empty(36, np.float32); AtB = np.empty(6, np.float32)
        cc = np.empty((self.n_corner, 4), np.float32); fc = np.empty(self.n_corner, np.uint8)
        cs = np.empty((self.n_surf, 4), np.float32); fs = np.empty(self.n_surf, np.uint8)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        n = self.L.ro_s2m_linearize(self.h, vp(tf), vp(AtA), vp(AtB), vp(cc), vp(fc), vp(cs), vp(fs))
        return {"AtA": AtA.reshape(6, 6), "AtB": AtB, "n_sel": n, "coeff_corner": cc, "flag_corner": fc, "coeff_surf": cs, "flag_surf": fs}

    def optimize(self, tf6, max_iter=30, edge_min=10, surf_min=100):
        tf = np.array(tf6, dtype=np.float32)
        conv, deg = C.c_int(), C.c_int()
        it = self.L.ro_s2m_optimize(self.h, C.c_void_p(tf.ctypes.data), max_iter, edge_min, surf_min, C.byref(conv), C.byref(deg))
        return {"pose": tf, "iterations": it, "converged": bool(conv.value), "degenerate": bool(deg.value)}


class SCManager:
    """Scan Context place recogniser (scancontext_oracle.cpp; src/scancontext/Scancontext.cpp)."""
    LIDAR_HEIGHT = 2.0

    def __init__(self):
        self.L = lib()
        self.L.ro_sc_create.restype = C.c_void_p
        self.h = C.c_void_p(self.L.ro_sc_create())

    def __del__(self):
        try:
            self.L.ro_sc_destroy(self.h)
        except Exception:
            pass

    def makeAndSaveScancontextAndKeys(self, scan):
        a = np.ascontiguousarray(scan, dtype=np.float32)
        self.L.ro_sc_make_and_save(self.h, C.c_void_p(a.ctypes.data), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]), C.c_double(self.LIDAR_HEIGHT))

    def size(self): return self.L.ro_sc_size(self.h)

    def get(self, index=-1):
        d = np.empty((20, 60), np.float64); rk = np.empty(20, np.float64); sk = np.empty(60, np.float64)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        assert self.L.ro_sc_get(self.h, index, vp(d), vp(rk), vp(sk)) == 0
        return d, rk, sk

    def distanceBtnScanContext(self, i, j):
        d = C.c_double(); s = C.c_int()
        assert self.L.ro_sc_distance(self.h, i, j, C.byref(d), C.byref(s)) == 0
        return d.value, s.value

    def detectLoopClosureID(self):
        yaw = C.c_float()
        lid = self.L.ro_sc_detect(self.h, C.byref(yaw))
        return lid, yaw.value


_REF_SC_PATH = os.path.join(_HERE, "_ref", "librolo_ref_sc.so")
_REF_FE_PATH = os.path.join(_HERE, "_ref", "librolo_ref_fe.so")
_REF_BM_PATH = os.path.join(_HERE, "_ref", "librolo_ref_bm.so")


def ref_sc_available() -> bool:
    return os.path.exists(_REF_SC_PATH)


class RefSCManager:
    """The REFERENCE's own SCManager (src/scancontext/Scancontext.cpp compiled from /root/reference with its vendored nanoflann,
    oracle/ref_sc_driver.cpp); same surface as SCManager above. Only for pinning the restated oracle."""

    def __init__(self):
        self.L = C.CDLL(_REF_SC_PATH)
        self.L.rsc_create.restype = C.c_void_p
        self.h = C.c_void_p(self.L.rsc_create())

    def __del__(self):
        try:
            self.L.rsc_destroy(self.h)
        except Exception:
            pass

    def makeAndSaveScancontextAndKeys(self, scan):
        a = np.ascontiguousarray(scan, dtype=np.float32)
        self.L.rsc_make_and_save(self.h, C.c_void_p(a.ctypes.data), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]))

    def size(self): return self.L.rsc_size(self.h)

    def get(self, index=-1):
        d = np.empty((20, 60), np.float64); rk = np.empty(20, np.float64); sk = np.empty(60, np.float64)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        assert self.L.rsc_get(self.h, index, vp(d), vp(rk), vp(sk)) == 0
        return d, rk, sk

    def distanceBtnScanContext(self, i, j):
        d = C.c_double(); s = C.c_int()
        assert self.L.rsc_distance(self.h, i, j, C.byref(d), C.byref(s)) == 0
        return d.value, s.value

    def detectLoopClosureID(self):
        yaw = C.c_float()
        lid = self.L.rsc_detect(self.h, C.byref(yaw))
        return lid, yaw.value


# ---- oracle/_ref: the REFERENCE's own rot_gicp sources compiled against oracle/ref_shim (see oracle/Makefile) ----
_REF_PATH = os.path.join(_HERE, "_ref", "librolo_ref.so")
_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


def build_ref() -> str | None:
    """Compiles the reference's rot_gicp headers from /root/reference (only possible where that checkout exists)."""
    if not os.path.isdir("/root/reference/include/rot_gicp"):
        return _REF_PATH if os.path.exists(_REF_PATH) else None
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return _REF_PATH


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(_REF_PATH)
        L.rr_create.restype = C.c_void_p
        for name in ("rr_linearize", "rr_so3_linearize", "rr_compute_error", "rr_t3_linearize", "rr_compute_t_error"):
            getattr(L, name).restype = C.c_double
        for name in ("rr_get_covariances", "rr_num_correspondences"):
            getattr(L, name).restype = C.c_size_t
        _ref = L
    return _ref


class Reference:
    """fast_gicp::RotVGICP<PointXYZI, PointXYZI> of the reference itself (oracle/ref_driver.cpp); same methods as Oracle."""

    def __init__(self):
        self.L = ref_lib()
        self.h = C.c_void_p(self.L.rr_create())

    def __del__(self):
        try:
            self.L.rr_destroy(self.h)
        except Exception:
            pass

    def setNumThreads(self, n): self.L.rr_set_num_threads(self.h, C.c_int(n))
    def setResolution(self, r): self.L.rr_set_resolution(self.h, C.c_double(r))
    def setPolarResolution(self, a, b, c): self.L.rr_set_polar_resolution(self.h, C.c_double(a), C.c_double(b), C.c_double(c))
    def setNeighborSearchMethod(self, m): self.L.rr_set_neighbor_search_method(self.h, C.c_int(m))
    def setRegularizationMethod(self, m): self.L.rr_set_regularization_method(self.h, C.c_int(m))
    def setCorrespondenceRandomness(self, k): self.L.rr_set_correspondence_randomness(self.h, C.c_int(k))
    def setOptimizerType(self, t): self.L.rr_set_optimizer_type(self.h, C.c_int(t))
    def setMaximumIterations(self, n): self.L.rr_set_max_iterations(self.h, C.c_int(n))
    def setRotationEpsilon(self, e): self.L.rr_set_rotation_epsilon(self.h, C.c_double(e))
    def setTransformationEpsilon(self, e): self.L.rr_set_transformation_epsilon(self.h, C.c_double(e))

    def setInputSource(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        self.L.rr_set_input_source(self.h, pts.ctypes.data_as(_fp), C.c_size_t(pts.shape[0]), C.c_size_t(pts.shape[1]))

    def setInputTarget(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        self.L.rr_set_input_target(self.h, pts.ctypes.data_as(_fp), C.c_size_t(pts.shape[0]), C.c_size_t(pts.shape[1]))

    def calculateCovariances(self, which): return self.L.rr_calculate_covariances(self.h, C.c_int(which))

    def getCovariances(self, which):
        n = self.L.rr_get_covariances(self.h, C.c_int(which), None)
        out = np.empty((n, 4, 4)); self.L.rr_get_covariances(self.h, C.c_int(which), _d(out)); return out

    def getSourceCovariances(self): return self.getCovariances(0)
    def getTargetCovariances(self): return self.getCovariances(1)

    def _set_covs(self, which, covs):
        covs = np.ascontiguousarray(covs, dtype=np.float64).reshape(-1, 16)
        self.L.rr_set_covariances(self.h, C.c_int(which), _d(covs), C.c_size_t(covs.shape[0]))

    def setSourceCovariances(self, c): self._set_covs(0, c)
    def setTargetCovariances(self, c): self._set_covs(1, c)

    def linearize(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16); H = np.empty((6, 6)); b = np.empty(6)
        return self.L.rr_linearize(self.h, _d(T), _d(H), _d(b)), H, b

    def so3_linearize(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16); H = np.empty((3, 3)); b = np.empty(3)
        return self.L.rr_so3_linearize(self.h, _d(T), _d(H), _d(b)), H, b

    def compute_error(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        return self.L.rr_compute_error(self.h, _d(T))

    def t3_linearize(self, t, g, l, dtn, dtn1):
        t, g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (t, g, l)); H = np.empty((6, 6)); b = np.empty(6)
        return self.L.rr_t3_linearize(self.h, _d(t), _d(g), _d(l), C.c_double(dtn), C.c_double(dtn1), _d(H), _d(b)), H, b

    def compute_t_error(self, t, g, l, dtn, dtn1):
        t, g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (t, g, l))
        return self.L.rr_compute_t_error(self.h, _d(t), _d(g), _d(l), C.c_double(dtn), C.c_double(dtn1))

    def setCtLambda(self, lam): self.L.rr_set_ct_lambda(self.h, C.c_float(lam))
    def numCorrespondences(self): return self.L.rr_num_correspondences(self.h)

    def align(self, guess=None):
        g = np.eye(4, dtype=np.float32) if guess is None else np.ascontiguousarray(guess, dtype=np.float32)
        out = np.empty((4, 4), dtype=np.float32)
        self.L.rr_align(self.h, g.ctypes.data_as(_fp), out.ctypes.data_as(_fp))
        self.final_transformation = out
        return out

    def getFinalPoseDouble(self):
        return self.final_transformation.astype(np.float64)   # the reference only keeps the float32 matrix

    def computeTranslation(self, trans, g, l, dtn, dtn1, lam):
        t = np.array(trans, dtype=np.float64); g, l = (np.ascontiguousarray(x, dtype=np.float64) for x in (g, l))
        self.L.rr_compute_translation(self.h, _d(t), _d(g), _d(l), C.c_double(dtn), C.c_double(dtn1), C.c_float(lam))
        return t

    def hasConverged(self): return bool(self.L.rr_has_converged(self.h))
    def nrIterations(self): return self.L.rr_nr_iterations(self.h)

    def getFinalHessian(self):
        H = np.empty((6, 6)); self.L.rr_get_final_hessian(self.h, _d(H)); return H

# ---- the reference's own range-image node sources (oracle/ref_fe_driver.cpp: src/imageProjection.cpp + src/featureExtraction.cpp compiled
# from /root/reference against oracle/ref_shim_ros): same call surface and result dictionaries as image_projection / feature_extraction
def ref_fe_available() -> bool:
    return os.path.exists(_REF_FE_PATH)


_ref_fe = {}


def _ref_fe_lib(horizon: int = 0):
    """projectPointCloud keeps `static float ang_res_x = 360.0 / float(Horizon_SCAN)` (imageProjection.cpp:439): initialised by the
    first call and never again, because a node runs with one Horizon_SCAN for its lifetime. A test process that projects frames of
    several widths therefore loads one private copy of the library per width (a copied file is a distinct shared object with its
    own statics)."""
    if horizon not in _ref_fe:
        import shutil, tempfile
        path = _REF_FE_PATH
        if _ref_fe:
            fd, path = tempfile.mkstemp(prefix=f"librolo_ref_fe_{horizon}_", suffix=".so"); os.close(fd)
            shutil.copyfile(_REF_FE_PATH, path)
        _ref_fe[horizon] = C.CDLL(path)
    return _ref_fe[horizon]


def ref_image_projection(raw: np.ndarray, n_scan, horizon, downsample_rate=1, min_range=2.0, max_range=1000.0, ring_offset=20, deskew=None):
    L = _ref_fe_lib(horizon)
    raw = np.ascontiguousarray(raw)
    cells = n_scan * horizon
    rim = np.empty((n_scan, horizon), dtype=np.float32)
    ext = np.empty((cells, 4), dtype=np.float32); col = np.empty(cells, dtype=np.int32); rng = np.empty(cells, dtype=np.float32)
    st = np.empty(n_scan, dtype=np.int32); en = np.empty(n_scan, dtype=np.int32)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    inc = np.ascontiguousarray(deskew["incre_rpy"] if deskew else (0, 0, 0), dtype=np.float32)
    n = L.rfe_image_projection(vp(raw), C.c_size_t(raw.shape[0]), C.c_int(raw.dtype.itemsize), C.c_int(ring_offset), C.c_int(n_scan), C.c_int(horizon),
                               C.c_int(downsample_rate), C.c_float(min_range), C.c_float(max_range), C.c_int(1 if deskew else 0),
                               C.c_int(deskew["time_offset"] if deskew else 0), vp(inc), C.c_float(deskew["scan_period"] if deskew else 0.1),
                               C.c_double(deskew["odom_time_diff"] if deskew else -1.0), vp(rim), vp(ext), vp(col), vp(rng), vp(st), vp(en))
    return {"range_image": rim, "extracted": ext[:n].copy(), "point_col_ind": col[:n].copy(), "point_range": rng[:n].copy(),
            "start_ring_index": st, "end_ring_index": en, "n": n}


def ref_feature_extraction(proj: dict, n_scan, horizon, edge_threshold=0.8, surf_threshold=0.1, surf_leaf_size=0.4):
    L = _ref_fe_lib(horizon)
    n = proj["n"]
    ext = np.ascontiguousarray(proj["extracted"], dtype=np.float32); col = np.ascontiguousarray(proj["point_col_ind"], dtype=np.int32)
    rng = np.ascontiguousarray(proj["point_range"], dtype=np.float32)
    st = np.ascontiguousarray(proj["start_ring_index"], dtype=np.int32); en = np.ascontiguousarray(proj["end_ring_index"], dtype=np.int32)
    curv = np.empty(n, dtype=np.float32); picked = np.empty(n, dtype=np.int32); label = np.empty(n, dtype=np.int32)
    corner = np.empty((n_scan * 120, 4), dtype=np.float32); surf = np.empty((n_scan * horizon, 4), dtype=np.float32)
    nc = C.c_int(); ns = C.c_int()
    vp = lambda a: C.c_void_p(a.ctypes.data)
    L.rfe_feature_extraction(vp(ext), vp(col), vp(rng), vp(st), vp(en), C.c_int(n), C.c_int(n_scan), C.c_int(horizon), C.c_float(edge_threshold),
                             C.c_float(surf_threshold), C.c_float(surf_leaf_size), vp(curv), vp(picked), vp(label), vp(corner), C.byref(nc), vp(surf), C.byref(ns))
    return {"curvature": curv, "neighbor_picked": picked, "label": label, "corner": corner[:nc.value].copy(), "surface": surf[:ns.value].copy()}

def ref_bm_available() -> bool:
    return os.path.exists(_REF_BM_PATH)


class RefScan2Map:
    """The REFERENCE's own scan-to-map member functions (cornerOptimization, surfOptimization, combineOptimizationCoeffs, LMOptimization,
    scan2MapOptimization of src/backMapping.cpp, compiled into oracle/_ref/librolo_ref_bm.so by oracle/ref_bm_driver.cpp). Same
    interface as Scan2Map where the two overlap."""

    def __init__(self, threads: int = 4):
        self.L = C.CDLL(_REF_BM_PATH)
        self.L.rbm_create.restype = C.c_void_p
        self.h = C.c_void_p(self.L.rbm_create())
        self.L.rbm_set_num_threads(self.h, threads)
        self.n_corner = self.n_surf = 0

    def __del__(self):
        try:
            self.L.rbm_destroy(self.h)
        except Exception:
            pass

    def setMap(self, corner, surf):
        c, cp, cn, cs = Scan2Map._c(corner); s, sp, sn, ss = Scan2Map._c(surf)
        assert cs == ss
        self.L.rbm_set_map(self.h, cp, cn, sp, sn, C.c_size_t(cs))

    def setScan(self, corner, surf):
        c, cp, cn, cs = Scan2Map._c(corner); s, sp, sn, ss = Scan2Map._c(surf)
        assert cs == ss
        self.n_corner, self.n_surf = c.shape[0], s.shape[0]
        self.L.rbm_set_scan(self.h, cp, cn, sp, sn, C.c_size_t(cs))

    def associate(self, tf6):
        """cornerOptimization + surfOptimization + combineOptimizationCoeffs at the pose: flags, coefficients, laserCloudSelNum."""
        tf = np.ascontiguousarray(tf6, dtype=np.float32)
        cc = np.empty((self.n_corner, 4), np.float32); fc = np.empty(self.n_corner, np.uint8)
        cs = np.empty((self.n_surf, 4), np.float32); fs = np.empty(self.n_surf, np.uint8)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        n = self.L.rbm_associate(self.h, vp(tf), vp(cc), vp(fc), vp(cs), vp(fs))
        return {"n_sel": n, "coeff_corner": cc, "flag_corner": fc, "coeff_surf": cs, "flag_surf": fs}

    def step(self, tf6, iter_count: int):
        """one pass of the loop body of scan2MapOptimization at the pose: returns the updated pose, LMOptimization's result, isDegenerate."""
        tf = np.array(tf6, dtype=np.float32)
        deg, sel = C.c_int(), C.c_int()
        conv = self.L.rbm_step(self.h, C.c_void_p(tf.ctypes.data), iter_count, C.byref(deg), C.byref(sel))
        return {"pose": tf, "converged": bool(conv), "degenerate": bool(deg.value), "n_sel": sel.value}

    def optimize(self, tf6, edge_min=10, surf_min=100):
        """scan2MapOptimization() itself (30 iterations at most)."""
        tf = np.array(tf6, dtype=np.float32)
        deg = C.c_int()
        rc = self.L.rbm_scan2map(self.h, C.c_void_p(tf.ctypes.data), edge_min, surf_min, C.byref(deg))
        return {"pose": tf, "enough_features": rc == 0, "degenerate": bool(deg.value)}


def require_ref(available, what: str):
    """Module-level guard of the tests that pin the oracle to oracle/_ref: where the reference checkout exists the library is built
    on demand and its absence is an ERROR (a silently skipped pin is no pin); elsewhere (the GPU box, where the prebuilt files
    travel with the tree) a missing library skips the module."""
    import pytest
    if available():
        return
    if os.path.isdir("/root/reference/include/rot_gicp"):
        build_ref()
        if not available():
            raise RuntimeError(f"{what} missing although /root/reference exists: `make -C oracle ref` failed")
        return
    pytest.skip(f"{what} not built and no reference checkout on this machine", allow_module_level=True)
