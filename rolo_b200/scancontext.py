"""Host-side mirror of the reference's SCManager (src/scancontext/Scancontext.cpp) over the C-ABI: the descriptor is built by a
CUDA kernel, matching runs in the library's host code."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class SCManager:
    LIDAR_HEIGHT = 2.0      # Scancontext.h:80

    def __init__(self, device: int = -1):
        self._L = capi.load()
        self._h = C.c_void_p()
        capi.check(self._L.rg_create(device, C.byref(self._h)))

    def __del__(self):
        try:
            self._L.rg_destroy(self._h)
        except Exception:
            pass

    def _ck(self, rc): return capi.check(rc, self._h)

    def makeAndSaveScancontextAndKeys(self, scan):
        a = np.ascontiguousarray(scan, dtype=np.float32)
        assert a.ndim == 2 and a.shape[1] >= 3
        self._ck(self._L.rg_sc_make_and_save(self._h, C.c_void_p(a.ctypes.data), a.shape[0], a.strides[0], self.LIDAR_HEIGHT))

    def size(self) -> int: return self._L.rg_sc_size(self._h)

    def get(self, index: int = -1):
        d = np.empty((20, 60), np.float64); rk = np.empty(20, np.float64); sk = np.empty(60, np.float64)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        self._ck(self._L.rg_sc_get(self._h, index, vp(d), vp(rk), vp(sk)))
        return d, rk, sk

    def distanceBtnScanContext(self, i: int, j: int):
        d = C.c_double(); s = C.c_int32()
        self._ck(self._L.rg_sc_distance(self._h, i, j, C.byref(d), C.byref(s)))
        return d.value, s.value

    def detectLoopClosureID(self):
        lid = C.c_int32(); yaw = C.c_float()
        self._ck(self._L.rg_sc_detect(self._h, C.byref(lid), C.byref(yaw)))
        return lid.value, yaw.value
