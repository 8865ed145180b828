"""Shared helpers for the parity tests: build an oracle and a CUDA registration with identical settings."""
import numpy as np

from oracle import binding as ob
from rolo_b200 import capi
from rolo_b200.registration import RotVGICP


def make_both(tgt, src, polar=(0.175, 0.175, 2.0), uniform=None, method=capi.DIRECT1, reg=capi.REG_PLANE, k=20,
              shared_covs=True):
    """Returns (oracle, gpu). With shared_covs the oracle's covariances are rounded to fp32 (the device
    storage format) and installed in BOTH, so every later comparison isolates the stage under test."""
    o = ob.Oracle()
    g = RotVGICP()
    for x in (o, g):
        if uniform is not None:
            x.setResolution(uniform)
        else:
            x.setPolarResolution(*polar)
        x.setNeighborSearchMethod(method)
        x.setRegularizationMethod(reg)
        x.setCorrespondenceRandomness(k)
        x.setInputTarget(tgt)
        x.setInputSource(src)
    if shared_covs:
        assert o.calculateCovariances(0) == 0 and o.calculateCovariances(1) == 0
        cs = o.getSourceCovariances().astype(np.float32).astype(np.float64)
        ct = o.getTargetCovariances().astype(np.float32).astype(np.float64)
        o.setSourceCovariances(cs); o.setTargetCovariances(ct)
        g.setSourceCovariances(cs); g.setTargetCovariances(ct)
    return o, g


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def oracle_features(n_scan, horizon):
    """raw frame -> (corner, surface) through the oracle's front end (for building scan-to-map inputs in CPU tests)."""
    def f(raw):
        p = ob.image_projection(raw, n_scan, horizon)
        q = ob.feature_extraction(p, n_scan, horizon)
        return q["corner"], q["surface"]
    return f
