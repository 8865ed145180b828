"""rolo/CloudInfoStamp at the ROS edge: the wire image packed from the device-resident front-end products
equals the oracle's byte for byte (both stages), and a serialised message installs the same feature cloud the reference's
consumer builds (lidarOdometry.cpp:521-524)."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP

pytestmark = pytest.mark.gpu


def _source_cloud(r: RotVGICP, n: int) -> np.ndarray:
    out = np.zeros((n, 4), dtype=np.float32)
    eye = np.eye(4, dtype=np.float32).reshape(16)
    capi.check(r._L.rg_transform_source(r._h, eye.ctypes.data_as(capi.fp), C.c_void_p(out.ctypes.data), out.strides[0]), r._h)
    return out


@pytest.mark.parametrize("sensor,rings,cols", [("small", 32, 512), ("os1-128", 128, 2048)])
def test_wire_image_is_bit_exact_and_round_trips(sensor, rings, cols):
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, synth.pose((0.5, -1.0, 20.0), (0.4, 0.2, 0.0)), sensor, seed=77, extra=500)
    p = ob.image_projection(raw, rings, cols); f = ob.feature_extraction(p, rings, cols)
    r = RotVGICP()
    nc, ns = r.frontEndFrame(raw, capi.TARGET, n_scan=rings, horizon_scan=cols)
    assert (nc, ns) == (len(f["corner"]), len(f["surface"])) and nc > 0 and ns > 0
    sc = np.linspace(-1, 1, 9).astype(np.float32); cov = np.arange(36, dtype=np.float32)
    msgs = []
    for stage in (0, 1):
        got = r.cloudInfoSerialize(stage, seq=9, stamp_sec=1234, stamp_nsec=5678, frame_id="os_sensor", scalars9=sc, covariance=cov, odom_available=True)
        want = ob.cloud_info_serialize(stage, 9, 1234, 5678, "os_sensor", p, f, sc, cov, True)
        assert len(got) == len(want)
        assert got == want, f"stage {stage}: first differing byte {next(i for i in range(len(got)) if got[i] != want[i])}"
        msgs.append(got)
    # consumer edge: a second context receives the stage-1 message
    c = RotVGICP()
    assert c.setInputFromCloudInfo(capi.SOURCE, msgs[1]) == (nc, ns)
    feat, onc, ons = ob.cloud_info_features(msgs[1])
    assert (onc, ons) == (nc, ns)
    cloud = _source_cloud(c, nc + ns)
    assert np.array_equal(cloud[:, :3], feat[:, :3])      # rg_transform_source writes xyz only
    # the stage-0 message carries no features: an empty cloud is installed, registration then refuses to run
    assert c.setInputFromCloudInfo(capi.SOURCE, msgs[0]) == (0, 0)
    with pytest.raises(Exception):
        c.setInputFromCloudInfo(capi.SOURCE, msgs[1][:-5])


def test_serialize_requires_a_resident_frame_and_capacity():
    r = RotVGICP()
    with pytest.raises(Exception):
        r.cloudInfoSerialize(0)
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, np.eye(4), "small", seed=5)
    r.frontEndFrame(raw, capi.TARGET, n_scan=32, horizon_scan=512)
    hdr = capi.MsgHeader(0, 0, 0, b"x"); n = C.c_size_t()
    assert r._L.rg_cloud_info_serialize(r._h, 1, C.byref(hdr), None, None, 0, C.byref(n)) == 0 and n.value > 0
    small = np.empty(16, dtype=np.uint8)
    assert r._L.rg_cloud_info_serialize(r._h, 1, C.byref(hdr), None, C.c_void_p(small.ctypes.data), small.size, C.byref(n)) == capi.RG_ERR_INVALID
