"""GPU parity of the two upstream range-image passes against the CPU oracle (bit-exact for index/integer work;
float outputs are produced by identical IEEE operations and must be equal too)."""
import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import synth
from rolo_b200.frontend import FrontEnd

pytestmark = pytest.mark.gpu

CASES = [("small", 32, 512), ("vlp64", 64, 1024), ("os1-128", 128, 2048)]


@pytest.mark.parametrize("sensor,n_scan,horizon", CASES)
def test_projection_and_features_match_oracle(sensor, n_scan, horizon):
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, synth.pose((1.0, -2.0, 30.0), (1.0, 0.5, 0.0)), sensor, extra=2000)
    fe = FrontEnd(n_scan=n_scan, horizon_scan=horizon)
    pg = fe.image_projection(raw)
    po = ob.image_projection(raw, n_scan, horizon)
    assert pg["n"] == po["n"] and pg["n"] > 0.9 * n_scan * horizon
    for k in ("range_image", "extracted", "point_col_ind", "point_range", "start_ring_index", "end_ring_index"):
        assert np.array_equal(pg[k], po[k]), k
    fg = fe.feature_extraction(po)
    fo = ob.feature_extraction(po, n_scan, horizon)
    assert np.array_equal(fg["curvature"], fo["curvature"])
    assert np.array_equal(fg["label"], fo["label"])
    assert np.array_equal(fg["neighbor_picked"], fo["neighbor_picked"])
    assert fg["corner"].shape == fo["corner"].shape and np.array_equal(fg["corner"], fo["corner"])
    assert fg["surface"].shape == fo["surface"].shape
    assert np.array_equal(fg["surface"], fo["surface"])
    assert len(fo["corner"]) > 0 and len(fo["surface"]) > 0


def test_projection_filters_and_first_writer_wins():
    """Hand-built frame: out-of-range points, bad rings, downsampling and two points in one cell."""
    from rolo_b200.frontend import RAW_POINT_DTYPE
    n_scan, horizon = 16, 256
    raw = np.zeros(8, dtype=RAW_POINT_DTYPE)
    xyz = [(10, 0, 0), (10.001, 0, 0), (1.0, 0, 0), (0, 10, 0), (0, -10, 1), (5, 5, 0), (2000, 0, 0), (3, 3, 3)]
    rings = [3, 3, 3, 20, 4, 5, 6, 7]
    for i, (p, r) in enumerate(zip(xyz, rings)):
        raw[i]["x"], raw[i]["y"], raw[i]["z"], raw[i]["ring"] = p[0], p[1], p[2], r
    fe = FrontEnd(n_scan=n_scan, horizon_scan=horizon, downsample_rate=1)
    pg = fe.image_projection(raw); po = ob.image_projection(raw, n_scan, horizon)
    assert pg["n"] == po["n"] == 4           # points 1 (same cell as 0), 2 (too near), 3 (ring 20), 6 (too far) are dropped
    for k in ("range_image", "extracted", "point_col_ind", "point_range", "start_ring_index", "end_ring_index"):
        assert np.array_equal(pg[k], po[k]), k
    assert pg["extracted"][0][0] == 10.0      # the first of the two competing points won
    fe2 = FrontEnd(n_scan=n_scan, horizon_scan=horizon, downsample_rate=2)
    assert fe2.image_projection(raw)["n"] == ob.image_projection(raw, n_scan, horizon, downsample_rate=2)["n"] == 1


def test_streamed_front_end_matches_oracle_chain():
    """BASELINE config 5 at small size: raw frame -> projection -> features -> `corner + surface` installed on the device
    -> SO3-LM align + translation LM; every frame's pose must match the CPU oracle running the same three stages."""
    from rolo_b200 import capi
    from rolo_b200.registration import RotVGICP
    sensor, rings, cols = "small", 32, 512
    scene = synth.make_scene()
    poses = [synth.pose((0.0, 0.0, 0.3 * k), (0.06 * k, 0.01 * k, 0.0)) for k in range(4)]
    raws = [synth.raw_frame(scene, poses[k], sensor, seed=100 + k) for k in range(4)]
    r = RotVGICP(); r.setPolarResolution(0.175, 0.175, 2.0)
    r.frontEndFrame(raws[0], capi.TARGET, n_scan=rings, horizon_scan=cols)
    prev = None
    for k in range(4):
        p = ob.image_projection(raws[k], rings, cols); f = ob.feature_extraction(p, rings, cols)
        cur = np.concatenate([f["corner"], f["surface"]], axis=0)
        if k > 0:
            r.swapSourceAndTarget()
            nc, ns = r.frontEndFrame(raws[k], capi.TARGET, n_scan=rings, horizon_scan=cols)
            assert (nc, ns) == (len(f["corner"]), len(f["surface"]))
            T = r.align()
            t = r.computeTranslation(np.zeros(3), T[:3, 3].astype(np.float64), np.zeros(3), 0.1, 0.1, 0.3)
            o = ob.Oracle(); o.setPolarResolution(0.175, 0.175, 2.0); o.setInputTarget(cur); o.setInputSource(prev)
            To = o.align()
            to = o.computeTranslation(np.zeros(3), To[:3, 3].astype(np.float64), np.zeros(3), 0.1, 0.1, 0.3)
            drot, _ = synth.pose_error(o.getFinalPoseDouble(), r.getFinalPoseDouble())
            assert drot <= 1e-4 and np.max(np.abs(t - to)) <= 1e-4
            assert r.nrIterations() == o.nrIterations()
        prev = cur


def test_deskew_matches_oracle():
    """deskewPoint (imageProjection.cpp:368-396) with deskewEnabled && odomAvailable: rotated points bit-equal to the oracle,
    range image / columns unchanged (they are computed before the correction)."""
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, synth.pose((1.0, -2.0, 30.0), (1.0, 0.5, 0.0)), "vlp64", extra=500)
    dk = {"time_offset": 24, "incre_rpy": np.array([0.004, -0.007, 0.031], np.float32), "scan_period": 0.1, "odom_time_diff": 0.0987}
    fe = FrontEnd(n_scan=64, horizon_scan=1024)
    plain = fe.image_projection(raw)
    fe.set_deskew(dk["incre_rpy"], dk["scan_period"], dk["odom_time_diff"], dk["time_offset"])
    pg = fe.image_projection(raw)
    po = ob.image_projection(raw, 64, 1024, deskew=dk)
    for k in ("range_image", "extracted", "point_col_ind", "point_range", "start_ring_index", "end_ring_index"):
        assert np.array_equal(pg[k], po[k]), k
    assert np.array_equal(pg["range_image"], plain["range_image"]) and np.array_equal(pg["extracted"][:, 3], plain["extracted"][:, 3])
    moved = np.linalg.norm(pg["extracted"][:, :3] - plain["extracted"][:, :3], axis=1)
    assert moved.max() > 0.05 and np.allclose(np.linalg.norm(pg["extracted"][:, :3], axis=1), np.linalg.norm(plain["extracted"][:, :3], axis=1), rtol=1e-5)
    fe.set_deskew(None)
    assert np.array_equal(fe.image_projection(raw)["extracted"], plain["extracted"])
