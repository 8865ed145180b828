"""GPU parity of the Scan Context recogniser (src/scancontext/Scancontext.cpp) against the CPU oracle through
the C-ABI: descriptor and keys bit-equal (a maximum per bin does not depend on point order; sums are sequential on both
sides), distances / shifts / loop ids equal."""
import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import synth
from rolo_b200.scancontext import SCManager

pytestmark = pytest.mark.gpu


def test_descriptor_keys_distance_and_detection_match_the_oracle():
    scene = synth.make_scene()
    o, g = ob.SCManager(), SCManager()
    poses = [synth.pose((0.5, -0.3, 3.0 * i), (-20.0 + 1.2 * i, -8.0 + 0.5 * i, 0.0)) for i in range(34)]
    poses.append(synth.pose((0, 0, 3.0 * 2 + 24.0), (-20.0 + 2.4, -8.0 + 1.0, 0.0)))
    for i, T in enumerate(poses):
        pts = synth.scan(scene, T, "small" if i % 3 else "vlp64", seed=50 + i)
        o.makeAndSaveScancontextAndKeys(pts); g.makeAndSaveScancontextAndKeys(pts)
        for a, b in zip(o.get(), g.get()):
            assert np.array_equal(a, b)
        if i in (5, 29, 30, 33, 34):
            assert o.detectLoopClosureID() == g.detectLoopClosureID()
    assert g.size() == o.size() == 35
    for i, j in ((0, 1), (34, 2), (10, 10), (3, 30)):
        assert o.distanceBtnScanContext(i, j) == g.distanceBtnScanContext(i, j)


def test_edge_cases():
    o, g = ob.SCManager(), SCManager()
    weird = np.array([[0, 0, 1, 0], [0, 5, -3, 0], [-5, 0, 2, 0], [0, -5, 2, 0], [5, 0, 2, 0], [79.9999, 0, 1, 0], [80.0001, 0, 9, 0], [300, 300, 50, 0],
                      [1e-30, 1e-30, 0.5, 0], [-1e-20, 3.0, -1002.0, 0], [3.0, -1e-20, -1002.0, 0], [np.nan, 1.0, 5.0, 0], [1.0, 1.0, np.nan, 0]], dtype=np.float32)
    for m in (o, g):
        m.makeAndSaveScancontextAndKeys(weird)
        m.makeAndSaveScancontextAndKeys(np.zeros((0, 4), np.float32))
    for idx in (0, 1):
        for a, b in zip(o.get(idx), g.get(idx)):
            assert np.array_equal(a, b, equal_nan=True)
    assert not g.get(1)[0].any()
    # a full-size frame: bit-equal as well
    pts = synth.scan(synth.make_scene(), np.eye(4), "os1-128", seed=1)
    o.makeAndSaveScancontextAndKeys(pts); g.makeAndSaveScancontextAndKeys(pts)
    assert np.array_equal(o.get()[0], g.get()[0])
