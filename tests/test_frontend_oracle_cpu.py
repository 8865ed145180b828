"""Independent (numpy) reading of what the front-end oracle must compute — its path is parity-unpinned by the reference (ROS node
code cannot be compiled here), so the restatement is cross-checked against a second, vectorised implementation written directly
from src/imageProjection.cpp:399-505 and src/featureExtraction.cpp:87-110."""
import numpy as np

from oracle import binding as ob
from rolo_b200 import synth


def _project_numpy(raw, n_scan, horizon, min_range=2.0, max_range=1000.0):
    x, y, z = raw["x"].astype(np.float32), raw["y"].astype(np.float32), raw["z"].astype(np.float32)
    ring = raw["ring"].astype(np.int64)
    rng = np.sqrt((x * x + y * y + z * z).astype(np.float32)).astype(np.float32)
    ang = (np.arctan2(x, y).astype(np.float32).astype(np.float64) * 180 / np.pi).astype(np.float32)        # :438
    res = np.float32(360.0 / np.float32(horizon))
    col = (-np.round((ang.astype(np.float64) - 90.0) / np.float64(res)) + horizon // 2).astype(np.int64)   # :441
    col = np.where(col >= horizon, col - horizon, col)
    ok = (rng >= min_range) & (rng <= max_range) & (ring >= 0) & (ring < n_scan) & (col >= 0) & (col < horizon)
    cell = ring * horizon + col
    first = {}
    for i in np.flatnonzero(ok):                       # first point to claim a cell wins (:451)
        first.setdefault(int(cell[i]), int(i))
    rim = np.full(n_scan * horizon, np.finfo(np.float32).max, np.float32)
    owner = np.full(n_scan * horizon, -1, np.int64)
    for c, i in first.items():
        rim[c] = rng[i]; owner[c] = i
    return rim.reshape(n_scan, horizon), owner.reshape(n_scan, horizon), (x, y, z, ring)


def test_projection_against_an_independent_implementation():
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, synth.pose((1.0, -2.0, 30.0), (1.0, 0.5, 0.0)), "small", extra=1500)
    po = ob.image_projection(raw, 32, 512)
    rim, owner, (x, y, z, ring) = _project_numpy(raw, 32, 512)
    # libm-vs-numpy atan2 may differ in the last bit: tolerate a handful of cells at column borders
    assert ((po["range_image"] != rim).sum()) <= 4
    same = po["range_image"] == rim
    occ = (owner >= 0) & same
    # row-major compaction (:477-505)
    idx = owner[occ]
    cols = np.nonzero(occ)[1]
    exp = np.stack([x[idx], y[idx], z[idx], (ring[idx].astype(np.float32) * z[idx]).astype(np.float32)], axis=1)
    if same.all():
        assert po["n"] == exp.shape[0] and np.array_equal(po["extracted"], exp) and np.array_equal(po["point_col_ind"], cols)
        assert np.array_equal(po["point_range"], rim[occ])
        counts = occ.sum(axis=1); ends = np.cumsum(counts)
        assert np.array_equal(po["start_ring_index"], ends - counts - 1 + 5) and np.array_equal(po["end_ring_index"], ends - 1 - 5)
    assert po["n"] > 0.9 * 32 * 512


def test_smoothness_against_an_independent_implementation():
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, synth.pose((0.0, 0.0, 10.0), (0.5, 0.5, 0.0)), "small")
    po = ob.image_projection(raw, 32, 512)
    fo = ob.feature_extraction(po, 32, 512)
    r = po["point_range"].astype(np.float64); n = r.shape[0]
    curv = np.zeros(n)
    for k in range(-5, 6):
        if k:
            curv[5:n - 5] += r[5 + k:n - 5 + k]
    curv[5:n - 5] -= 10 * r[5:n - 5]
    curv = curv ** 2                                           # featureExtraction.cpp:92-101
    got = fo["curvature"].astype(np.float64)
    assert np.allclose(got[5:n - 5], curv[5:n - 5], rtol=2e-3, atol=1e-6)
    # corners are picked among high-curvature points, surfaces among low-curvature ones (:189-243)
    corner_xyz = {tuple(p[:3]) for p in fo["corner"]}
    ext = po["extracted"]
    is_corner = np.array([tuple(p[:3]) in corner_xyz for p in ext])
    assert is_corner.sum() == fo["corner"].shape[0] and got[is_corner].min() > 0.8 * (1 - 1e-6)
