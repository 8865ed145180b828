"""rolo/CloudInfoStamp wire image: the oracle's serialiser is checked against an INDEPENDENT parser written
straight from the message definition (msg/CloudInfoStamp.msg:1-28, sensor_msgs/PointCloud2, std_msgs/Header) and against a
committed golden digest. No GPU."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cloud_info_digest.json")


class Rd:
    def __init__(self, b): self.b, self.o = b, 0
    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o); self.o += struct.calcsize("<" + fmt); return v if len(v) > 1 else v[0]
    def string(self):
        n = self.take("I"); s = self.b[self.o:self.o + n].decode(); self.o += n; return s
    def array(self, dt):
        n = self.take("I"); a = np.frombuffer(self.b, dtype=dt, count=n, offset=self.o); self.o += n * np.dtype(dt).itemsize; return a
    def header(self): return {"seq": self.take("I"), "stamp": self.take("II"), "frame_id": self.string()}
    def cloud(self):
        c = {"header": self.header(), "height": self.take("I"), "width": self.take("I")}
        c["fields"] = [(self.string(), self.take("I"), self.take("B"), self.take("I")) for _ in range(self.take("I"))]
        c["is_bigendian"], c["point_step"], c["row_step"] = self.take("B"), self.take("I"), self.take("I")
        c["data"] = self.array(np.uint8); c["is_dense"] = self.take("B")
        return c


def parse(msg: bytes) -> dict:
    r = Rd(msg)
    m = {"header": r.header(), "startRingIndex": r.array("<i4"), "endRingIndex": r.array("<i4"), "pointColInd": r.array("<i4"), "pointRange": r.array("<f4")}
    m["scalars"] = np.array(r.take("9f"), dtype=np.float32)
    m["covariance"] = r.array("<f4"); m["odomAvailable"] = r.take("B")
    for name in ("cloud_projected", "extracted_corner", "extracted_surface", "extracted_normal", "extracted_ground"):
        m[name] = r.cloud()
    assert r.o == len(msg)
    return m


def xyzi(c: dict) -> np.ndarray:
    rec = c["data"].view(np.float32).reshape(-1, 8) if c["data"].size else np.zeros((0, 8), np.float32)
    return rec[:, [0, 1, 2, 4]]


def fake_products(seed=3, n_scan=4, n=37, nc=5, ns=11):
    g = np.random.default_rng(seed)
    proj = {"start_ring_index": g.integers(0, 50, n_scan).astype(np.int32), "end_ring_index": g.integers(0, 50, n_scan).astype(np.int32),
            "point_col_ind": g.integers(0, 2048, n).astype(np.int32), "point_range": g.random(n, dtype=np.float32) * 50,
            "extracted": g.standard_normal((n, 4)).astype(np.float32), "n": n}
    feat = {"corner": g.standard_normal((nc, 4)).astype(np.float32), "surface": g.standard_normal((ns, 4)).astype(np.float32)}
    return proj, feat


@pytest.mark.parametrize("stage", [0, 1])
def test_oracle_wire_image_matches_the_message_definition(stage):
    proj, feat = fake_products()
    sc = np.arange(9, dtype=np.float32) * 0.25
    cov = np.linspace(0, 1, 36, dtype=np.float32)
    msg = ob.cloud_info_serialize(stage, 17, 1700000000, 123456789, "velodyne", proj, feat, sc, cov, True)
    m = parse(msg)
    assert m["header"] == {"seq": 17, "stamp": (1700000000, 123456789), "frame_id": "velodyne"}
    assert np.array_equal(m["scalars"], sc) and np.array_equal(m["covariance"], cov) and m["odomAvailable"] == 1
    if stage == 0:
        for k, name in (("startRingIndex", "start_ring_index"), ("endRingIndex", "end_ring_index"), ("pointColInd", "point_col_ind"), ("pointRange", "point_range")):
            assert np.array_equal(m[k], proj[name])
    else:   # featureExtraction.cpp:268-274 freeCloudInfoMemory
        assert all(m[k].size == 0 for k in ("startRingIndex", "endRingIndex", "pointColInd", "pointRange"))
    p = m["cloud_projected"]
    assert p["header"] == {"seq": 0, "stamp": (1700000000, 123456789), "frame_id": "velodyne"}
    assert (p["height"], p["width"], p["point_step"], p["row_step"], p["is_bigendian"], p["is_dense"]) == (1, 37, 32, 32 * 37, 0, 1)
    assert p["fields"] == [("x", 0, 7, 1), ("y", 4, 7, 1), ("z", 8, 7, 1), ("intensity", 16, 7, 1)]
    assert np.array_equal(xyzi(p), proj["extracted"])
    rec = p["data"].view(np.float32).reshape(-1, 8)
    assert np.all(rec[:, 3] == 1.0) and np.all(rec[:, 5:] == 0.0)
    blank = {"header": {"seq": 0, "stamp": (0, 0), "frame_id": ""}, "height": 0, "width": 0, "fields": [], "is_bigendian": 0, "point_step": 0, "row_step": 0, "is_dense": 0}
    def is_blank(c): return all(c[k] == v for k, v in blank.items()) and c["data"].size == 0
    assert is_blank(m["extracted_ground"])
    if stage == 0:
        assert is_blank(m["extracted_corner"]) and is_blank(m["extracted_surface"]) and is_blank(m["extracted_normal"])
    else:
        assert np.array_equal(xyzi(m["extracted_corner"]), feat["corner"]) and np.array_equal(xyzi(m["extracted_surface"]), feat["surface"])
        nrm = m["extracted_normal"]
        assert (nrm["height"], nrm["width"], nrm["point_step"], nrm["is_dense"], len(nrm["fields"])) == (1, 0, 32, 1, 4) and nrm["data"].size == 0


def test_consumer_side_restatement():
    proj, feat = fake_products(seed=9, nc=0, ns=7)
    msg = ob.cloud_info_serialize(1, 0, 1, 2, "lidar", proj, feat)
    cloud, nc, ns = ob.cloud_info_features(msg)
    assert (nc, ns) == (0, 7) and np.array_equal(cloud, feat["surface"])
    proj, feat = fake_products(seed=10)
    cloud, nc, ns = ob.cloud_info_features(ob.cloud_info_serialize(1, 0, 1, 2, "lidar", proj, feat))
    assert np.array_equal(cloud, np.vstack([feat["corner"], feat["surface"]]))
    with pytest.raises(ValueError):
        ob.cloud_info_features(msg[:-3])


def test_golden_digest():
    """tests/golden/cloud_info_digest.json was produced by tools/make_cloud_info_golden.py (same call as below)."""
    want = json.load(open(GOLD))
    for stage in (0, 1):
        proj, feat = fake_products(seed=21 + stage, n_scan=8, n=301, nc=40, ns=77)
        msg = ob.cloud_info_serialize(stage, 5, 100, 200, "base_link", proj, feat, np.arange(9, dtype=np.float32), np.ones(36, np.float32), stage == 1)
        assert len(msg) == want[f"stage{stage}"]["size"] and hashlib.sha256(msg).hexdigest() == want[f"stage{stage}"]["sha256"]
