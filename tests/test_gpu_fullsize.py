"""GPU parity at the BENCHMARKED sizes (BASELINE configs[1] and the shape of configs[3]).

* configs[1]: two 128 x 2048 OS1-128 clouds, POLAR(0.175, 0.175, 2.0) / DIRECT1 / PLANE / k = 20 — exact kNN sets on ALL points,
  covariances, per-call {H, b, cost} and the pose of a 20-iteration Gauss-Newton align() against the reference's own sources
  (oracle/_ref) when they were built, else the oracle port.
* configs[3] shape: UNIFORM 1.0 m voxels / DIRECT7, a >= 1 M-point target (8 scans merged in the map frame) and a 131 k-point source.

Tolerances as in test_gpu_parity.py: index work bit-exact; covariances 3e-7 (fp32 storage); per-call sums 1e-9 relative with the same
fp32-rounded covariances on both sides; pose <= 1e-4 rad / 1e-4 m (north_star).
"""
import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP

from helpers import make_both, rel_err

pytestmark = pytest.mark.gpu

POSE_TOL_RAD = 1e-4
POSE_TOL_M = 1e-4


def _cpu_reference():
    """The reference's own rot_gicp sources when oracle/_ref exists (it travels with the tree), else the oracle port."""
    return ob.Reference() if ob.ref_available() else ob.Oracle()


def _check_covs(a, b):
    scale = np.maximum(1.0, np.abs(a).max(axis=(1, 2)))
    assert np.max(np.abs(a - b).max(axis=(1, 2)) / scale) <= 3e-7


@pytest.fixture(scope="module")
def c2_pair():
    return synth.scan_pair("os1-128")        # BASELINE configs[1]: 128 x 2048 = 262144 points per cloud


def test_c2_knn_and_covariances_all_points(c2_pair):
    tgt, src, _ = c2_pair
    assert tgt.shape[0] == 128 * 2048
    o, g = make_both(tgt, src, shared_covs=False)
    a = o.knn(1, 20); b = g.knn(capi.TARGET, 20)
    assert np.array_equal(a, b), f"{np.sum(np.any(a != b, axis=1))} of {a.shape[0]} neighbour lists differ"
    assert o.calculateCovariances(0) == 0 and o.calculateCovariances(1) == 0
    g.calculate_covariances(capi.SOURCE); g.calculate_covariances(capi.TARGET)
    _check_covs(o.getSourceCovariances(), g.getSourceCovariances())
    _check_covs(o.getTargetCovariances(), g.getTargetCovariances())


def test_c2_per_call_sums(c2_pair):
    tgt, src, T_gt = c2_pair
    o, g = make_both(tgt, src)               # identical fp32-rounded covariances on both sides
    for T in (np.eye(4), T_gt):
        co, Ho, bo = o.linearize(T)
        cg, Hg, bg = g.linearize(T)
        assert g.numCorrespondences() == len(o.getCorrespondences()[0])
        assert abs(cg - co) <= 1e-9 * abs(co)
        assert rel_err(Hg, Ho) <= 1e-9 and rel_err(bg, bo) <= 1e-9
        assert abs(g.compute_error(T) - o.compute_error(T)) <= 1e-9 * abs(co)
    co, Ho, bo = o.so3_linearize(T_gt); cg, Hg, bg = g.so3_linearize(T_gt)
    assert rel_err(Hg, Ho) <= 1e-9 and rel_err(bg, bo) <= 1e-9 and abs(cg - co) <= 1e-9 * abs(co)


def test_c2_align_20_gauss_newton_iterations(c2_pair):
    tgt, src, T_gt = c2_pair
    ref = _cpu_reference()
    g = RotVGICP()
    for x in (ref, g):
        x.setPolarResolution(0.175, 0.175, 2.0)
        x.setOptimizerType(capi.OPT_GAUSS_NEWTON); x.setMaximumIterations(20)
        x.setRotationEpsilon(1e-30); x.setTransformationEpsilon(1e-30)
        x.setInputTarget(tgt); x.setInputSource(src)
    ref.align(); Tg = g.align()
    assert g.nrIterations() == ref.nrIterations() == 19
    drot, dtr = synth.pose_error(ref.getFinalPoseDouble(), g.getFinalPoseDouble())
    assert drot <= POSE_TOL_RAD and dtr <= POSE_TOL_M, (drot, dtr)
    erot, etr = synth.pose_error(Tg.astype(np.float64), T_gt)
    assert erot < 2e-3 and etr < 5e-3


@pytest.fixture(scope="module")
def submap_case():
    """configs[3] shape at 1/10 scale: 8 scans of 128 x 1024 merged in the map frame (1.05 M points) + a live scan near the first pose."""
    scene = synth.make_scene()
    n_scans, sensor = 8, "os1-128x1024"
    clouds, poses = [], []
    for k in range(n_scans):
        a = 2 * np.pi * k / 77
        T = synth.pose((0.0, 0.0, np.degrees(a) + 90.0), (5.0 + 18.0 * np.cos(a), 3.0 + 10.0 * np.sin(a), 0.0))
        p = synth.scan(scene, T, sensor, sigma=0.004, seed=1000 + k)
        q = p.copy(); q[:, :3] = (p[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        clouds.append(q); poses.append(T)
    target = np.ascontiguousarray(np.concatenate(clouds, axis=0))
    T_live = poses[0] @ synth.pose((1.0, -1.5, 2.0), (0.10, -0.05, 0.02))
    source = synth.scan(scene, T_live, sensor, seed=7)
    return target, source, poses[0], T_live


def test_c4_shape_uniform_direct7(submap_case):
    target, source, guess, T_live = submap_case
    assert target.shape[0] >= 1_000_000
    o, g = make_both(target, source, uniform=1.0, method=capi.DIRECT7, shared_covs=False)
    # covariances of the 1 M-point target (and of the source): exact kNN underneath
    assert o.calculateCovariances(1) == 0 and o.calculateCovariances(0) == 0
    g.calculate_covariances(capi.TARGET); g.calculate_covariances(capi.SOURCE)
    ct = o.getTargetCovariances(); cs = o.getSourceCovariances()
    _check_covs(ct, g.getTargetCovariances()); _check_covs(cs, g.getSourceCovariances())
    # same fp32-rounded covariances on both sides, then voxel map + per-call sums
    ct = ct.astype(np.float32).astype(np.float64); cs = cs.astype(np.float32).astype(np.float64)
    for x in (o, g):
        x.setTargetCovariances(ct); x.setSourceCovariances(cs)
    o.createVoxelmap(); g.create_voxelmap()
    ko, mo, co_, no = o.getVoxels(); kg, mg, cg_, ng = g.getVoxels()
    assert np.array_equal(ko, kg) and np.array_equal(no, ng)
    assert np.max(np.abs(mo[:, :3] - mg)) <= 1e-9
    for T in (guess, T_live):
        co, Ho, bo = o.linearize(T); cg, Hg, bg = g.linearize(T)
        assert g.numCorrespondences() == len(o.getCorrespondences()[0])
        assert abs(cg - co) <= 1e-9 * abs(co) and rel_err(Hg, Ho) <= 1e-9 and rel_err(bg, bo) <= 1e-9
    for x in (o, g):
        x.setOptimizerType(capi.OPT_GAUSS_NEWTON); x.setMaximumIterations(10)
        x.setRotationEpsilon(1e-30); x.setTransformationEpsilon(1e-30)
    o.align(guess.astype(np.float32)); Tg = g.align(guess.astype(np.float32))
    drot, dtr = synth.pose_error(o.getFinalPoseDouble(), g.getFinalPoseDouble())
    assert drot <= POSE_TOL_RAD and dtr <= POSE_TOL_M, (drot, dtr)
    erot, etr = synth.pose_error(Tg.astype(np.float64), T_live)
    assert erot < 2e-3 and etr < 1e-2
