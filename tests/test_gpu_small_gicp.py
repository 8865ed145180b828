"""The small_gicp-shaped adapter (include/small_gicp, SURVEY.md §8 f3 second half): the reference's commented VGICP prior-patch sequence
(src/backMapping.cpp:1993-2028) compiled against it and run on the GPU core; the result must equal the same registration configured
through the rot_gicp surface (UNIFORM leaf, DIRECT1, k = 10, PLANE, 6-DoF LM, 50 iterations) and agree with the CPU oracle.
small_gicp itself is not in the reference tree (no pinned version): there is no numerical parity target beyond this mapping."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEAF = 1.0


def _build(tmp_path):
    libdir = os.path.dirname(capi.lib_path())
    exe = tmp_path / "small_gicp_sequence"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-w", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "ref_shim"),
                           os.path.join(ROOT, "tests", "cpp", "small_gicp_sequence.cpp"), "-o", str(exe), "-L", libdir, "-lrolo_gicp_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_adapter_compiles_cpu(tmp_path):
    capi.load()
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_adapter_sequence_matches_the_core_and_the_oracle(tmp_path):
    exe = _build(tmp_path)
    tgt, src, T_gt = synth.scan_pair("small")
    tgt.tofile(tmp_path / "t.bin"); src.tofile(tmp_path / "s.bin")
    out = subprocess.run([str(exe), str(tmp_path / "t.bin"), str(tmp_path / "s.bin"), str(LEAF)], capture_output=True, text=True, check=True).stdout
    T = np.array([float(x) for x in out.splitlines()[0].split()[1:]]).reshape(4, 4)
    m = re.search(r"error=(\S+) converged=(\d) iterations=(\d+) inliers=(\d+) cov00=(\S+)", out)
    err, conv, iters, inliers, cov00 = float(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4)), float(m.group(5))

    def configure(r):
        r.setResolution(LEAF); r.setNeighborSearchMethod(capi.DIRECT1); r.setCorrespondenceRandomness(10)
        r.setOptimizerType(capi.OPT_LM); r.setMaximumIterations(50)
        r.setRotationEpsilon(0.1 * np.pi / 180.0); r.setTransformationEpsilon(1e-3)
        r.setInputTarget(tgt); r.setInputSource(src)
    g = RotVGICP(); configure(g); g.align()
    o = ob.Oracle(); configure(o); o.align()
    # the adapter IS the core with that configuration
    assert np.max(np.abs(T - g.getFinalPoseDouble())) <= 1e-9
    assert iters == g.nrIterations() + 1 and conv == int(g.hasConverged()) and inliers == g.numCorrespondences()
    assert abs(cov00 - g.getSourceCovariances()[0, 0, 0]) <= 1e-9          # printed with 10 significant digits
    # and the core agrees with the CPU oracle (north_star tolerance), and registers the pair
    drot, dtr = synth.pose_error(o.getFinalPoseDouble(), T)
    assert drot <= 1e-4 and dtr <= 1e-4, (drot, dtr)
    erot, etr = synth.pose_error(T, T_gt)
    assert erot < 5e-3 and etr < 5e-2
    assert err > 0 and inliers > 0.5 * len(src)
