"""Pins the restated Scan Context oracle against the REFERENCE ITSELF: oracle/_ref/librolo_ref_sc.so is the reference's own
src/scancontext/Scancontext.cpp (+ its vendored nanoflann kd-tree) compiled from /root/reference against oracle/ref_shim
(`make -C oracle ref`). Descriptors, keys, pairwise distances / shifts and loop detections must agree exactly."""
import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import synth

ob.require_ref(ob.ref_sc_available, "oracle/_ref/librolo_ref_sc.so")


def test_oracle_equals_the_reference_scancontext():
    scene = synth.make_scene()
    o, r = ob.SCManager(), ob.RefSCManager()
    poses = [synth.pose((0.5, -0.3, 3.0 * i), (-20.0 + 1.2 * i, -8.0 + 0.5 * i, 0.0)) for i in range(44)]
    poses.append(synth.pose((0, 0, 3.0 * 2 + 24.0), (-20.0 + 2.4, -8.0 + 1.0, 0.0)))      # revisit of pose 2 with another heading
    found = 0
    for i, T in enumerate(poses):
        pts = synth.scan(scene, T, "small" if i % 3 else "vlp64", seed=50 + i)
        o.makeAndSaveScancontextAndKeys(pts); r.makeAndSaveScancontextAndKeys(pts)
        for a, b in zip(o.get(), r.get()):
            assert np.array_equal(a, b), i
        do, dr = o.detectLoopClosureID(), r.detectLoopClosureID()      # includes the reference's every-10th-call tree rebuild
        assert do == dr, (i, do, dr)
        found += do[0] >= 0
    assert found >= 3 and o.size() == r.size() == 45
    for i, j in ((0, 1), (44, 2), (10, 10), (3, 30), (40, 7), (17, 44)):
        assert o.distanceBtnScanContext(i, j) == r.distanceBtnScanContext(i, j)


def test_corner_cases_agree():
    o, r = ob.SCManager(), ob.RefSCManager()
    weird = np.array([[0, 5, -3, 0], [-5, 0, 2, 0], [0, -5, 2, 0], [5, 0, 2, 0], [79.9999, 0, 1, 0], [80.0001, 0, 9, 0], [300, 300, 50, 0],
                      [1e-30, 1e-30, 0.5, 0], [-1e-20, 3.0, -1002.0, 0], [3.0, -1e-20, -1002.0, 0]], dtype=np.float32)
    for m in (o, r):
        m.makeAndSaveScancontextAndKeys(weird)
    for a, b in zip(o.get(), r.get()):
        assert np.array_equal(a, b)
