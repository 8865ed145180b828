"""Host-side mirror of the two upstream range-image passes over the C-ABI.

image_projection  == ImageProjection::projectPointCloud + cloudExtraction (src/imageProjection.cpp:399-505)
feature_extraction == FeatureExtraction::calculateSmoothness/markOccludedPoints/extractFeatures
                      (src/featureExtraction.cpp:87-266); outputs mirror rolo/CloudInfoStamp
                      (msg/CloudInfoStamp.msg: startRingIndex, endRingIndex, pointColInd, pointRange, clouds).
All numerics run in librolo_gicp_b200.so (CUDA); this module only marshals arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

#: byte layout of the reference's VelodynePointXYZIRT (imageProjection.cpp:7-19): x y z pad | intensity ring pad time pad
RAW_POINT_DTYPE = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                            "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                            "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})


def _ctx():
    h = C.c_void_p()
    capi.check(capi.load().rg_create(-1, C.byref(h)))
    return h


class FrontEnd:
    """Owns one C-ABI context (device scratch for the range image and feature buffers)."""

    def __init__(self, n_scan=128, horizon_scan=2048, downsample_rate=1, lidar_min_range=2.0, lidar_max_range=1000.0,
                 edge_threshold=0.8, surf_threshold=0.1, surf_leaf_size=0.4):
        self.L = capi.load()
        self.h = _ctx()
        self.n_scan, self.horizon = n_scan, horizon_scan
        self.pp = capi.ProjectionParams(n_scan, horizon_scan, downsample_rate, lidar_min_range, lidar_max_range, RAW_POINT_DTYPE.itemsize, 20)
        self.fp = capi.FeatureParams(n_scan, horizon_scan, edge_threshold, surf_threshold, surf_leaf_size)

    def __del__(self):
        try:
            self.L.rg_destroy(self.h)
        except Exception:
            pass

    def set_deskew(self, incre_rpy=None, scan_period: float = 0.1, odom_time_diff: float = 0.1, time_offset: int = 24):
        """deskewPoint (imageProjection.cpp:368-396): rotation-only correction by the odometry increment; None disables it."""
        if incre_rpy is None:
            capi.check(self.L.rg_set_deskew(self.h, None), self.h)
            return
        p = capi.DeskewParams(time_offset, float(incre_rpy[0]), float(incre_rpy[1]), float(incre_rpy[2]), scan_period, odom_time_diff)
        capi.check(self.L.rg_set_deskew(self.h, C.byref(p)), self.h)

    def image_projection(self, raw: np.ndarray) -> dict:
        raw = np.ascontiguousarray(raw)
        assert raw.dtype.itemsize == self.pp.point_stride_bytes
        cells = self.n_scan * self.horizon
        rim = np.empty((self.n_scan, self.horizon), dtype=np.float32)
        ext = np.empty((cells, 4), dtype=np.float32); col = np.empty(cells, dtype=np.int32); rng = np.empty(cells, dtype=np.float32)
        st = np.empty(self.n_scan, dtype=np.int32); en = np.empty(self.n_scan, dtype=np.int32); n = C.c_int32()
        vp = lambda a: C.c_void_p(a.ctypes.data)
        capi.check(self.L.rg_image_projection(self.h, vp(raw), C.c_size_t(raw.shape[0]), C.byref(self.pp), vp(rim), vp(ext), vp(col), vp(rng),
                                              vp(st), vp(en), C.byref(n)), self.h)
        k = n.value
        return {"range_image": rim, "extracted": ext[:k].copy(), "point_col_ind": col[:k].copy(), "point_range": rng[:k].copy(),
                "start_ring_index": st, "end_ring_index": en, "n": k}

    def feature_extraction(self, proj: dict) -> dict:
        n = proj["n"]
        ext = np.ascontiguousarray(proj["extracted"], dtype=np.float32); col = np.ascontiguousarray(proj["point_col_ind"], dtype=np.int32)
        rng = np.ascontiguousarray(proj["point_range"], dtype=np.float32)
        st = np.ascontiguousarray(proj["start_ring_index"], dtype=np.int32); en = np.ascontiguousarray(proj["end_ring_index"], dtype=np.int32)
        curv = np.empty(n, dtype=np.float32); picked = np.empty(n, dtype=np.int32); label = np.empty(n, dtype=np.int32)
        corner = np.empty((self.n_scan * 120, 4), dtype=np.float32); surf = np.empty((self.n_scan * self.horizon, 4), dtype=np.float32)
        nc = C.c_int32(); ns = C.c_int32()
        vp = lambda a: C.c_void_p(a.ctypes.data)
        capi.check(self.L.rg_feature_extraction(self.h, vp(ext), vp(col), vp(rng), vp(st), vp(en), C.c_int32(n), C.byref(self.fp), vp(curv), vp(picked),
                                                vp(label), vp(corner), C.byref(nc), vp(surf), C.byref(ns)), self.h)
        return {"curvature": curv, "neighbor_picked": picked, "label": label, "corner": corner[:nc.value].copy(), "surface": surf[:ns.value].copy()}
