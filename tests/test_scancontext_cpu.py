"""Scan Context oracle (src/scancontext/Scancontext.cpp) without a GPU: the descriptor against a numpy
restatement of the bin rule, the yaw-shift property of the descriptor, and loop detection on a synthetic revisit."""
import numpy as np

from oracle import binding as ob
from rolo_b200 import synth


def _desc_numpy(pts, h=2.0):
    x, y, z = pts[:, 0].astype(np.float32), pts[:, 1].astype(np.float32), (pts[:, 2].astype(np.float64) + h).astype(np.float32)
    rng = np.sqrt(x * x + y * y).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        th = np.degrees(np.arctan2(y.astype(np.float64), x.astype(np.float64))) % 360.0
    d = np.full((20, 60), -1000.0)
    ok = rng <= 80.0
    ring = np.clip(np.ceil(rng.astype(np.float64) / 80.0 * 20).astype(int), 1, 20) - 1
    sec = np.clip(np.ceil(th / 360.0 * 60).astype(int), 1, 60) - 1
    np.maximum.at(d, (ring[ok], sec[ok]), z[ok].astype(np.float64))
    d[d == -1000.0] = 0.0
    return d


def test_descriptor_matches_the_bin_rule_and_keys_are_means():
    scene = synth.make_scene()
    pts = synth.scan(scene, synth.pose((0, 0, 20.0), (1.0, 2.0, 0.0)), "vlp64", seed=4)
    m = ob.SCManager(); m.makeAndSaveScancontextAndKeys(pts)
    d, rk, sk = m.get()
    ref = _desc_numpy(pts)
    # arctan2-based angles agree with the reference's quadrant formula except within rounding of a bin edge
    assert (d != ref).sum() <= 2
    assert np.allclose(rk, d.mean(axis=1), rtol=1e-14) and np.allclose(sk, d.mean(axis=0), rtol=1e-14)
    assert (d > 0).sum() > 300


def test_yaw_rotation_shifts_columns_and_distance_recovers_it():
    scene = synth.make_scene()
    a = synth.scan(scene, synth.pose((0, 0, 0.0), (2.0, 1.0, 0.0)), "vlp64", seed=4, sigma=0.0)
    k = 7                                   # 7 sectors = 42 degrees
    b = synth.scan(scene, synth.pose((0, 0, -6.0 * k), (2.0, 1.0, 0.0)), "vlp64", seed=4, sigma=0.0)
    m = ob.SCManager(); m.makeAndSaveScancontextAndKeys(a); m.makeAndSaveScancontextAndKeys(b)
    dist, shift = m.distanceBtnScanContext(0, 1)
    assert dist < 0.15 and shift in (k, 60 - k)
    d0, sh0 = m.distanceBtnScanContext(0, 0)
    assert abs(d0) < 1e-12 and sh0 == 0


def test_loop_detection_on_a_revisit():
    scene = synth.make_scene()
    m = ob.SCManager()
    poses = [synth.pose((0, 0, 3.0 * i), (-20.0 + 1.2 * i, -8.0 + 0.5 * i, 0.0)) for i in range(34)]
    for i, T in enumerate(poses):
        m.makeAndSaveScancontextAndKeys(synth.scan(scene, T, "small", seed=50 + i))
        if i < 30:
            assert m.detectLoopClosureID() == (-1, 0.0)           # fewer than NUM_EXCLUDE_RECENT + 1 entries
    # revisit pose 2 with a different heading
    m.makeAndSaveScancontextAndKeys(synth.scan(scene, synth.pose((0, 0, 3.0 * 2 + 24.0), (-20.0 + 2.4, -8.0 + 1.0, 0.0)), "small", seed=999))
    lid, yaw = m.detectLoopClosureID()
    yaw_deg = float(np.degrees(yaw))
    assert lid in (1, 2, 3) and min(abs(yaw_deg - 24.0), abs(360.0 - yaw_deg - 24.0)) <= 6.0 + 1e-3
