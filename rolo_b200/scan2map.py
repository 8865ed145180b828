"""Host-side mirror of backMapping's scan-to-map optimisation over the C-ABI (src/backMapping.cpp:681-1054).

Method names follow the reference's member functions; all numerics run in librolo_gicp_b200.so (CUDA kernel per
Gauss-Newton iteration + the 6x6 host algebra). Poses are the reference's transformTobeMapped: roll pitch yaw x y z."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class Scan2Map:
    def __init__(self, device: int = -1):
        self._L = capi.load()
        self._h = C.c_void_p()
        capi.check(self._L.rg_create(device, C.byref(self._h)))
        self._keep = []
        self.n_corner = self.n_surf = 0

    def __del__(self):
        try:
            self._L.rg_destroy(self._h)
        except Exception:
            pass

    def _ck(self, rc): return capi.check(rc, self._h)

    @staticmethod
    def _c(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        assert a.ndim == 2 and a.shape[1] >= 3
        return a

    def setMap(self, corner, surf):
        """kdtreeCornerFromMap / kdtreeSurfFromMap ->setInputCloud (backMapping.cpp:689-690)."""
        c, s = self._c(corner), self._c(surf)
        assert c.shape[1] == s.shape[1]
        self._ck(self._L.rg_s2m_set_map(self._h, C.c_void_p(c.ctypes.data), c.shape[0], C.c_void_p(s.ctypes.data), s.shape[0], c.strides[0]))

    def setScan(self, corner, surf):
        c, s = self._c(corner), self._c(surf)
        assert c.shape[1] == s.shape[1]
        self.n_corner, self.n_surf = c.shape[0], s.shape[0]
        self._ck(self._L.rg_s2m_set_scan(self._h, C.c_void_p(c.ctypes.data), c.shape[0], C.c_void_p(s.ctypes.data), s.shape[0], c.strides[0]))

    def linearize(self, tf6) -> dict:
        """cornerOptimization + surfOptimization + combineOptimizationCoeffs + AtA/AtB at one pose (backMapping.cpp:719-1006)."""
        tf = np.ascontiguousarray(tf6, dtype=np.float32)
        AtA = np.empty(36, np.float32); AtB = np.empty(6, np.float32); n = C.c_int32()
        vp = lambda a: C.c_void_p(a.ctypes.data)
        self._ck(self._L.rg_s2m_linearize(self._h, vp(tf), vp(AtA), vp(AtB), C.byref(n)))
        cc = np.empty((self.n_corner, 4), np.float32); fc = np.empty(self.n_corner, np.uint8)
        cs = np.empty((self.n_surf, 4), np.float32); fs = np.empty(self.n_surf, np.uint8)
        self._ck(self._L.rg_s2m_get_coefficients(self._h, vp(cc), vp(fc), vp(cs), vp(fs)))
        return {"AtA": AtA.reshape(6, 6), "AtB": AtB, "n_sel": n.value, "coeff_corner": cc, "flag_corner": fc, "coeff_surf": cs, "flag_surf": fs}

    def optimize(self, tf6, max_iter: int = 30, edge_min: int = 10, surf_min: int = 100) -> dict:
        """scan2MapOptimization (backMapping.cpp:681-710) without transformUpdate()."""
        tf = np.array(tf6, dtype=np.float32)
        res = capi.S2MResult()
        self._ck(self._L.rg_s2m_optimize(self._h, C.c_void_p(tf.ctypes.data), max_iter, edge_min, surf_min, C.byref(res)))
        return {"pose": tf, "iterations": res.iterations, "converged": bool(res.converged), "degenerate": bool(res.degenerate),
                "enough_features": bool(res.enough_features)}
