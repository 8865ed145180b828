"""The plain-C example (examples/lidar_odometry_sequence.c) links against the C-ABI only and reproduces the Python API."""
import os
import re
import subprocess

import numpy as np
import pytest

from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_example_matches_python(tmp_path):
    libdir = os.path.dirname(capi.lib_path())
    exe = tmp_path / "lidar_odometry_sequence"
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "lidar_odometry_sequence.c"),
                           "-o", str(exe), "-L", libdir, "-lrolo_gicp_b200", f"-Wl,-rpath,{libdir}", "-lm"])
    tgt, src, _ = synth.scan_pair("small")
    tgt.tofile(tmp_path / "target.bin"); src.tofile(tmp_path / "source.bin")
    out = subprocess.run([str(exe), str(tmp_path / "target.bin"), str(tmp_path / "source.bin")], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("frame")]
    assert len(lines) == 3
    r = RotVGICP(); r.setPolarResolution(0.175, 0.175, 2.0); r.setInputTarget(tgt); r.setInputSource(src)
    T = r.align()
    t = r.computeTranslation(np.zeros(3), T[:3, 3].astype(np.float64), np.zeros(3), 0.1, 0.1, 0.3)
    for l in lines:
        nums = [float(x) for x in re.findall(r"-?\d+\.\d+", l)]
        assert np.allclose(nums[:3], T[0, :3], atol=2e-6) and np.allclose(nums[3:6], t, atol=2e-6)
