"""The product's C++ drop-in header (include/rot_gicp/gicp/rot_vgicp.hpp) compiled and run with the reference caller's own
statement sequence (lidarOdometry.cpp:453-500). PCL/Eigen are absent from the image, so the test compiles against
oracle/ref_shim (test infrastructure); a ROLO checkout uses its own PCL/Eigen."""
import os
import re
import subprocess

import numpy as np
import pytest

from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    libdir = os.path.dirname(capi.lib_path())
    exe = tmp_path / "drop_in_sequence"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-w", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "ref_shim"),
                           os.path.join(ROOT, "tests", "cpp", "drop_in_sequence.cpp"), "-o", str(exe), "-L", libdir, "-lrolo_gicp_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_drop_in_header_compiles_cpu(tmp_path):
    capi.load()
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_drop_in_sequence_matches_python(tmp_path):
    exe = _build(tmp_path)
    tgt, src, _ = synth.scan_pair("small")
    tgt.tofile(tmp_path / "t.bin"); src.tofile(tmp_path / "s.bin")
    out = subprocess.run([str(exe), str(tmp_path / "t.bin"), str(tmp_path / "s.bin")], capture_output=True, text=True, check=True).stdout
    r = RotVGICP(); r.setPolarResolution(0.175, 0.175, 2.0); r.setInputTarget(tgt); r.setInputSource(src)
    T = r.align()
    t = r.computeTranslation(np.zeros(3), T[:3, 3].astype(np.float64), np.array([0.01, 0.0, 0.0]), 0.1, 0.1, 0.3)
    nums = [float(x) for x in re.findall(r"-?\d+\.\d+", out.splitlines()[0])]
    assert "converged=1" in out and "alias=invalid_argument" in out and f"covs={len(src)}" in out
    assert np.allclose(nums[:3], T[0, :3], atol=2e-6) and np.allclose(nums[3:6], t, atol=2e-6)
    assert np.allclose(nums[6:9], src[0, :3] + t.astype(np.float32), atol=1e-4)
    m = re.search(r"subclass converged=(\d) err=(\S+)", out)     # a user subclass of fast_gicp::LsqRegistration through the inherited align()
    assert m and m.group(1) == "1" and float(m.group(2)) < 1e-4
    fit = float(re.search(r"fitness=(\S+)", out).group(1))
    assert abs(fit - r.getFitnessScore(T)) <= 1e-9 * fit      # output cloud = input + translation (lsq.hpp:75-78)
