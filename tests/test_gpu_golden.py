"""GPU: the CUDA path (through the C-ABI) against the committed golden fixtures — no oracle call, no /root/reference."""
import os

import numpy as np
import pytest

from rolo_b200 import capi, synth
from rolo_b200.frontend import FrontEnd
from rolo_b200.registration import RotVGICP

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _reg(tgt, src):
    r = RotVGICP(); r.setPolarResolution(0.175, 0.175, 2.0); r.setInputTarget(tgt); r.setInputSource(src)
    return r


def test_registration_matches_golden():
    g = np.load(os.path.join(G, "registration_tiny.npz"))
    tgt, src, T_gt = synth.scan_pair("tiny")
    for name, opt, iters, eps in (("gn", capi.OPT_GAUSS_NEWTON, 6, 1e-30), ("lm", capi.OPT_LM, None, None), ("so3lm", capi.OPT_SO3_LM, None, None)):
        r = _reg(tgt, src); r.setOptimizerType(opt)
        if iters:
            r.setMaximumIterations(iters); r.setRotationEpsilon(eps); r.setTransformationEpsilon(eps)
        T = r.align()
        drot, dtr = synth.pose_error(r.getFinalPoseDouble(), g[f"{name}_final"])
        assert drot <= 1e-4 and dtr <= 1e-4                                     # north_star bar
        assert r.nrIterations() == int(g[f"{name}_iters"]) and r.hasConverged() == bool(g[f"{name}_converged"])
        if name == "so3lm":
            t = r.computeTranslation(np.zeros(3), T[:3, 3].astype(np.float64), np.array([0.01, 0.0, 0.0]), 0.1, 0.1, 0.3)
            assert np.max(np.abs(t - g["so3lm_translation"])) <= 1e-4
    r = _reg(tgt, src)
    assert np.array_equal(r.knn(capi.TARGET, 20)[:64], g["knn_target_first64"])     # index work: bit-exact
    r.calculate_covariances(capi.SOURCE)
    assert np.max(np.abs(r.getSourceCovariances()[:64] - g["cov_source_first64"])) <= 3e-7
    r.create_voxelmap()
    k, m, c6, n = r.getVoxels()
    assert np.array_equal(k, g["vox_keys"]) and np.array_equal(n, g["vox_npts"])
    assert np.max(np.abs(m - g["vox_mean"][:, :3])) <= 1e-9
    assert np.max(np.abs((c6[:, 0] + c6[:, 3] + c6[:, 5]) - g["vox_cov_trace"])) <= 1e-6
    for method, tag in ((capi.DIRECT1, "d1"), (capi.DIRECT7, "d7")):
        r.setNeighborSearchMethod(method)
        c, H, b = r.linearize(T_gt)
        # covariances here are computed on each side (fp32 storage on the device): 1e-6 relative
        assert abs(c - float(g[f"lin_{tag}_cost"])) <= 1e-6 * abs(float(g[f"lin_{tag}_cost"]))
        assert np.max(np.abs(H - g[f"lin_{tag}_H"])) <= 1e-6 * np.max(np.abs(g[f"lin_{tag}_H"]))
        assert np.max(np.abs(b - g[f"lin_{tag}_b"])) <= 1e-6 * np.max(np.abs(g[f"lin_{tag}_H"]))


def test_frontend_matches_golden():
    g = np.load(os.path.join(G, "frontend_small.npz"))
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, synth.pose((1.0, -2.0, 30.0), (1.0, 0.5, 0.0)), "small", extra=300)
    fe = FrontEnd(n_scan=32, horizon_scan=512)
    p = fe.image_projection(raw)
    f = fe.feature_extraction(p)
    assert p["n"] == int(g["n"]) and np.array_equal(p["start_ring_index"], g["start_ring_index"]) and np.array_equal(p["end_ring_index"], g["end_ring_index"])
    assert np.array_equal(p["point_col_ind"][:256], g["point_col_ind_head"]) and np.array_equal(p["point_range"][:256], g["point_range_head"])
    assert np.array_equal(f["curvature"][:256], g["curvature_head"])
    assert np.array_equal(np.bincount(f["label"] + 1, minlength=3), g["label_hist"])
    assert np.array_equal(f["corner"], g["corner"]) and len(f["surface"]) == int(g["n_surface"]) and np.array_equal(f["surface"][:128], g["surface_head"])
