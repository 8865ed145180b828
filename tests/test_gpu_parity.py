"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Tolerances (written here on purpose):
  * integer/index work (kNN sets, voxel keys, counts, correspondences): bit-exact.
  * covariances: 3e-7 relative to max(1, largest entry of the matrix) (device storage is fp32).
  * voxel means 1e-9 m, voxel covariances 1e-6.
  * per-call H / b / cost with identical fp32-rounded covariances on both sides: 1e-9 relative
    (fp64 accumulate; only summation order differs).
  * final pose: north_star's bar, <= 1e-4 rad and <= 1e-4 m against the oracle.
"""
import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP, RotVGICPBatch

from helpers import make_both, rel_err
from conftest import poses_for_tests

pytestmark = pytest.mark.gpu

POSE_TOL_RAD = 1e-4
POSE_TOL_M = 1e-4


def test_knn_exact(small_pair):
    tgt, src, _ = small_pair
    o, g = make_both(tgt, src, shared_covs=False)
    for which in (0, 1):
        for k in (20, 10, 7):
            a = o.knn(which, k)
            b = g.knn(which, k)
            assert np.array_equal(a, b), f"kNN mismatch which={which} k={k}: {np.sum(np.any(a != b, axis=1))} rows differ"


@pytest.mark.parametrize("reg", [capi.REG_PLANE, capi.REG_NONE, capi.REG_MIN_EIG, capi.REG_NORMALIZED_MIN_EIG, capi.REG_FROBENIUS, capi.REG_PLANE_S])
def test_covariances(small_pair, reg):
    tgt, src, _ = small_pair
    o, g = make_both(tgt, src, reg=reg, shared_covs=False)
    assert o.calculateCovariances(0) == 0
    g.calculate_covariances(capi.SOURCE)
    a = o.getSourceCovariances(); b = g.getSourceCovariances()
    assert a.shape == b.shape
    # device storage is fp32: error is relative to the magnitude of each matrix (FROBENIUS yields entries ~1e3)
    scale = np.maximum(1.0, np.abs(a).max(axis=(1, 2)))
    assert np.max(np.abs(a - b).max(axis=(1, 2)) / scale) <= 3e-7
    assert np.all(b[:, 3, :] == 0) and np.all(b[:, :, 3] == 0)


@pytest.mark.parametrize("grid", ["polar", "uniform"])
def test_voxelmap(small_pair, grid):
    tgt, src, _ = small_pair
    o, g = make_both(tgt, src, uniform=1.0 if grid == "uniform" else None)
    o.createVoxelmap(); g.create_voxelmap()
    ko, mo, co, no = o.getVoxels()
    kg, mg, cg, ng = g.getVoxels()
    assert np.array_equal(ko, kg) and np.array_equal(no, ng)
    assert np.max(np.abs(mo[:, :3] - mg)) <= 1e-9
    co6 = np.stack([co[:, 0, 0], co[:, 0, 1], co[:, 0, 2], co[:, 1, 1], co[:, 1, 2], co[:, 2, 2]], axis=1)
    assert np.max(np.abs(co6 - cg)) <= 1e-9     # same fp32-rounded inputs on both sides -> fp64 sums agree


@pytest.mark.parametrize("grid", ["polar", "uniform"])
def test_deterministic_voxelmap_is_bit_reproducible(small_pair, grid):
    """rg_set_deterministic: points sorted by voxel key, every voxel summed in point-index order = the summation order of the reference's
    serial loop (vmp_voxel.hpp:167-197). The map is then identical from run to run AND equal to the oracle's bit for bit; registrations on
    it take the same optimiser decisions every time."""
    tgt, src, _ = small_pair
    o, g = make_both(tgt, src, uniform=1.0 if grid == "uniform" else None)
    g.setDeterministic(True)
    o.createVoxelmap(); g.create_voxelmap()
    ko, mo, co, no = o.getVoxels()
    kg, mg, cg, ng = g.getVoxels()
    assert np.array_equal(ko, kg) and np.array_equal(no, ng)
    co6 = np.stack([co[:, 0, 0], co[:, 0, 1], co[:, 0, 2], co[:, 1, 1], co[:, 1, 2], co[:, 2, 2]], axis=1)
    assert np.array_equal(mo[:, :3], mg) and np.array_equal(co6, cg)            # bit for bit
    g.create_voxelmap()
    k2, m2, c2, n2 = g.getVoxels()
    assert np.array_equal(mg, m2) and np.array_equal(cg, c2)
    runs = []
    for _ in range(2):
        r = RotVGICP(); r.setDeterministic(True)
        if grid == "uniform":
            r.setResolution(1.0)
        else:
            r.setPolarResolution(0.175, 0.175, 2.0)
        r.setOptimizerType(capi.OPT_LM); r.setInputTarget(tgt); r.setInputSource(src)
        r.align(); runs.append((r.getFinalPoseDouble().copy(), r.nrIterations(), r.getFinalHessian().copy()))
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1] and np.array_equal(runs[0][2], runs[1][2])


@pytest.mark.parametrize("grid,method", [("polar", capi.DIRECT1), ("polar", capi.DIRECT7), ("uniform", capi.DIRECT1),
                                         ("uniform", capi.DIRECT7), ("uniform", capi.DIRECT27)])
def test_linearize_per_call(small_pair, grid, method):
    tgt, src, _ = small_pair
    o, g = make_both(tgt, src, uniform=1.0 if grid == "uniform" else None, method=method)
    noff = {capi.DIRECT1: 1, capi.DIRECT7: 7, capi.DIRECT27: 27}[method]
    for T in poses_for_tests():
        co, Ho, bo = o.linearize(T)
        cg, Hg, bg = g.linearize(T)
        # correspondences: same (point, voxel) set, bit-exact keys
        idx, key = o.getCorrespondences()
        kg = g.getCorrespondenceKeys(noff).reshape(-1, noff, 3)
        valid = kg[:, :, 0] != np.iinfo(np.int32).min
        assert int(valid.sum()) == len(idx) == g.numCorrespondences()
        got = {(int(i), *map(int, kg[i, j])) for i, j in zip(*np.nonzero(valid))}
        exp = {(int(i), *map(int, k)) for i, k in zip(idx, key)}
        assert got == exp
        assert rel_err(cg, co) <= 1e-9
        assert rel_err(Hg, Ho) <= 1e-9 and rel_err(bg, bo) <= 1e-9
        assert np.allclose(Hg, Hg.T, rtol=0, atol=1e-9 * np.abs(Hg).max())
        # SO(3) system == top-left block of the SE(3) one (rot_vgicp_impl.hpp:345-366)
        c3, H3, b3 = g.so3_linearize(T)
        _, H3o, b3o = o.so3_linearize(T)
        assert rel_err(H3, H3o) <= 1e-9 and rel_err(b3, b3o) <= 1e-9 and rel_err(c3, co) <= 1e-9
        assert rel_err(H3, Hg[:3, :3]) <= 1e-12 and rel_err(b3, bg[:3]) <= 1e-12
        # LM trial evaluations reuse correspondences + Mahalanobis of the linearisation pose
        for dT in (synth.pose((0.1, 0.2, -0.1), (0.01, 0.0, -0.02)), synth.pose((-0.4, 0.0, 0.3), (0.0, 0.03, 0.0))):
            Tt = dT @ T
            assert rel_err(g.compute_error(Tt), o.compute_error(Tt)) <= 1e-9
        # cost-only evaluation (evaluateCost with H == nullptr)
        cc, _, _ = g.linearize(T, want_H=False)
        assert rel_err(cc, co) <= 1e-9


def test_translation_stage_per_call(small_pair):
    tgt, src, _ = small_pair
    o, g = make_both(tgt, src)
    T = synth.pose((2.0, -3.0, 5.0), (0.0, 0.0, 0.0))
    o.so3_linearize(T); g.so3_linearize(T)
    o.setCtLambda(0.3); g.setCtLambda(0.3)
    guess = np.array([0.2, -0.1, 0.05]); last = np.array([0.18, -0.12, 0.04])
    for t in (np.zeros(3), np.array([0.01, -0.02, 0.005]), np.array([-0.3, 0.1, 0.2])):
        co, Ho, bo = o.t3_linearize(t, guess, last, 0.1, 0.1)
        cg, Hg, bg = g.t3_linearize(t, guess, last, 0.1, 0.1)
        assert rel_err(cg, co) <= 1e-9 and rel_err(Hg, Ho) <= 1e-9 and rel_err(bg, bo) <= 1e-9
        assert rel_err(g.compute_t_error(t, guess, last, 0.1, 0.1), o.compute_t_error(t, guess, last, 0.1, 0.1)) <= 1e-9


def _configure(x, opt, max_iter=None, eps=None):
    x.setOptimizerType(opt)
    if max_iter is not None:
        x.setMaximumIterations(max_iter)
    if eps is not None:
        x.setRotationEpsilon(eps); x.setTransformationEpsilon(eps)


@pytest.mark.parametrize("opt,max_iter,eps", [(capi.OPT_GAUSS_NEWTON, 20, 1e-30), (capi.OPT_LM, None, None), (capi.OPT_SO3_LM, None, None)])
def test_align_matches_oracle(small_pair, opt, max_iter, eps):
    """Full align(): covariances computed on each side (GPU kNN+SVD vs oracle), then the optimiser."""
    tgt, src, T_gt = small_pair
    o, g = make_both(tgt, src, shared_covs=False)
    _configure(o, opt, max_iter, eps); _configure(g, opt, max_iter, eps)
    To = o.align(); Tg = g.align()
    drot, dtrans = synth.pose_error(o.getFinalPoseDouble(), g.getFinalPoseDouble())
    assert drot <= POSE_TOL_RAD and dtrans <= POSE_TOL_M, (drot, dtrans)
    assert g.nrIterations() == o.nrIterations()
    assert g.hasConverged() == o.hasConverged()
    assert np.max(np.abs(To - Tg)) <= 1e-4
    if opt != capi.OPT_SO3_LM:
        assert rel_err(g.getFinalHessian(), o.getFinalHessian()) <= 1e-4
        # and it actually registers the pair
        erot, etr = synth.pose_error(Tg.astype(np.float64), T_gt)
        assert erot < 5e-3 and etr < 2e-2


def test_front_end_sequence(small_pair):
    """The exact call sequence of LidarOdometry::scanRegeistration (lidarOdometry.cpp:460-494)."""
    tgt, src, _ = small_pair
    res = []
    for make in (ob.Oracle, RotVGICP):
        r = make()
        r.setPolarResolution(0.175, 0.175, 2.0)
        r.setNumThreads(0)
        r.setInputTarget(tgt)
        r.setInputSource(src)
        T = r.align()
        t = r.computeTranslation(np.zeros(3), T[:3, 3].astype(np.float64), np.array([0.01, 0.0, 0.0]), 0.1, 0.1, 0.3)
        res.append((r.getFinalPoseDouble(), t))
    drot, _ = synth.pose_error(res[0][0], res[1][0])
    assert drot <= POSE_TOL_RAD
    assert np.max(np.abs(res[0][1] - res[1][1])) <= POSE_TOL_M


def test_batch_equals_single(tiny_pair):
    tgt, src, T_gt = tiny_pair
    rng = np.random.default_rng(99)
    scene = synth.make_scene()
    pairs = []
    for i in range(5):
        T = synth.random_offset(rng)
        pairs.append((synth.scan(scene, np.eye(4), "tiny", seed=1234 + i), synth.scan(scene, T, "tiny", seed=5678 + i), T))
    batch = RotVGICPBatch(len(pairs))
    singles = []
    for i, (t, s, _) in enumerate(pairs):
        for r in (batch[i],):
            r.setPolarResolution(0.175, 0.175, 2.0); r.setOptimizerType(capi.OPT_LM)
            r.setInputTarget(t); r.setInputSource(s)
        r1 = RotVGICP()
        r1.setPolarResolution(0.175, 0.175, 2.0); r1.setOptimizerType(capi.OPT_LM)
        r1.setInputTarget(t); r1.setInputSource(s)
        singles.append(r1.align().copy())
    out = batch.align()
    for i in range(len(pairs)):
        assert np.max(np.abs(out[i] - singles[i])) <= 1e-6
    assert batch.lastLinearizeLaunches() > 0


def test_error_behaviour(tiny_pair):
    tgt, src, _ = tiny_pair
    r = RotVGICP()
    r.setInputTarget(tgt); r.setInputSource(src)
    with pytest.raises(ValueError):          # polar resolution never set: reference would divide by zero (Q4)
        r.align()
    r.setPolarResolution(0.175, 0.175, 2.0)
    out = src.copy()
    r.align(output=out)
    with pytest.raises(ValueError):          # rot_vgicp_impl.hpp:150-152
        r.align(output=r._keep[capi.SOURCE])
    r.setNeighborSearchMethod(capi.DIRECT_RADIUS)
    with pytest.raises(capi.RoloError):      # reference abort()s (vmp_voxel.hpp:16-18)
        r.align()
    # transformed output cloud = T * source (lsq.hpp:178)
    r.setNeighborSearchMethod(capi.DIRECT1)
    T = r.align(output=out)
    exp = (src[:, :3].astype(np.float64) @ T[:3, :3].astype(np.float64).T + T[:3, 3]).astype(np.float32)
    assert np.max(np.abs(out[:, :3] - exp)) <= 1e-4


def test_c1_full_size_properties(c1_pair):
    """BASELINE config 1 (64 x 1024): pose vs oracle, and domain properties at full size."""
    tgt, src, T_gt = c1_pair
    o, g = make_both(tgt, src, shared_covs=False)
    for x in (o, g):
        _configure(x, capi.OPT_GAUSS_NEWTON, 20, 1e-30)
    o.align(); Tg = g.align()
    drot, dtrans = synth.pose_error(o.getFinalPoseDouble(), g.getFinalPoseDouble())
    assert drot <= POSE_TOL_RAD and dtrans <= POSE_TOL_M
    erot, etr = synth.pose_error(Tg.astype(np.float64), T_gt)
    assert erot < 2e-3 and etr < 5e-3
    # identity registration: source == target -> the gradient is tiny against |H| (each point sits in the voxel
    # whose mean it contributed to) and the pose stays at identity
    r = RotVGICP(); r.setPolarResolution(0.175, 0.175, 2.0); r.setOptimizerType(capi.OPT_GAUSS_NEWTON); r.setMaximumIterations(3)
    r.setInputTarget(tgt); r.setInputSource(tgt.copy())
    c, H, b = r.linearize(np.eye(4))
    assert np.linalg.norm(b) <= 1e-4 * np.linalg.norm(H)
    assert r.numCorrespondences() == tgt.shape[0]        # every point finds its own voxel
    Ti = r.align()
    erot, etr = synth.pose_error(Ti.astype(np.float64), np.eye(4))
    assert erot < 2e-4 and etr < 5e-3      # voxel-mean attraction moves an identical cloud by ~2 mm (reference behaviour)
