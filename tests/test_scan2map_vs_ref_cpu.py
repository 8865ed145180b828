"""Pins oracle/scan2map_oracle.cpp (the restated scan-to-map optimisation) against the REFERENCE'S OWN member functions
of class mapOptimization (src/backMapping.cpp: pointAssociateToMap, trans2Affine3f, updatePointAssociateToMap, cornerOptimization,
surfOptimization, combineOptimizationCoeffs, LMOptimization, scan2MapOptimization), compiled from /root/reference into
oracle/_ref/librolo_ref_bm.so (oracle/ref_bm_driver.cpp; the member functions are taken from the reference file at build time by
oracle/extract_ref_members.py and run unmodified, with the reference's include/rolo/utility.h in front). Where /root/reference exists the
library MUST exist (the tests fail, they do not skip); elsewhere (the GPU box) the prebuilt library travels with the tree.

Both sides call the same restated third-party numerics (oracle/s2m_numerics.hpp: cv::eigen, cv::solve, cv::Mat::inv, cv::gemm,
Eigen::ColPivHouseholderQR — themselves pinned against real OpenCV in test_scan2map_cpu.py) and the same exact kd-tree, so every
difference would be a difference in the transcription of the REFERENCE's code: float / double promotions, thresholds, the order the
selected rows are stacked in, the degeneracy projection, the convergence test. Equality is bit for bit.

The reference runs with numberOfCores = 1 here. Its masks laserCloudOriCornerFlag / laserCloudOriSurfFlag are bit-packed std::vector<bool>
(backMapping.cpp:136,139) written from `#pragma omp parallel for` loops (:722,833): two threads whose index ranges meet inside one 64-bit
word can lose each other's update, which drops a selected point now and then (observed here: poses differing in the 6th digit from run
to run with 4 threads). That is a data race of the reference, not a definition; the oracle and the CUDA path implement the race-free
meaning, which is what one thread computes.
"""
import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import synth
import helpers

ob.require_ref(ob.ref_bm_available, "oracle/_ref/librolo_ref_bm.so")


def _pair(case, threads=4):
    o = ob.Scan2Map(threads=threads); r = ob.RefScan2Map(threads=1)
    for x in (o, r):
        x.setMap(case["map_corner"], case["map_surf"]); x.setScan(case["scan_corner"], case["scan_surf"])
    return o, r


@pytest.fixture(scope="module")
def small_case():
    return synth.scan2map_case(helpers.oracle_features(32, 512), "small")


@pytest.fixture(scope="module")
def vlp_case():
    return synth.scan2map_case(helpers.oracle_features(64, 1024), "vlp64", n_keyframes=4, seed=23, guess_deg=2.0, guess_t=0.3)


def _same_association(o, r, tf):
    a = o.linearize(tf); b = r.associate(tf)
    assert a["n_sel"] == b["n_sel"]
    for key in ("flag_corner", "flag_surf", "coeff_corner", "coeff_surf"):
        assert np.array_equal(a[key], b[key]), key
    return a["n_sel"]


def test_selection_and_coefficients_equal_the_reference(small_case, vlp_case):
    for case in (small_case, vlp_case):
        o, r = _pair(case)
        n = _same_association(o, r, case["pose_guess"])
        assert n > 1000
        _same_association(o, r, case["pose_true"])
        far = case["pose_guess"].copy(); far[3:] += np.float32(0.8)          # many points beyond the 1 m neighbour gate / the s > 0.1 gate
        assert _same_association(o, r, far) < n


def test_every_gauss_newton_step_equals_the_reference(small_case, vlp_case):
    """The loop body of scan2MapOptimization one pass at a time: both sides start each pass from the same pose (the reference's), so a
    difference cannot hide behind the next pass. The pose after the pass, LMOptimization's return value and isDegenerate must be equal."""
    for case in (small_case, vlp_case):
        o, r = _pair(case)
        tf = case["pose_guess"].copy()
        seen_convergence = False
        for it in range(30):
            ref = r.step(tf, it)
            # the oracle's optimise() with max_iter = 1 is exactly one pass, but it treats every pass as iterCount == 0; so compare the
            # first pass directly and the later ones through a full run below
            if it == 0:
                mine = o.optimize(tf, max_iter=1)
                assert np.array_equal(mine["pose"], ref["pose"]) and mine["converged"] == ref["converged"] and mine["degenerate"] == ref["degenerate"]
            tf = ref["pose"]
            if ref["converged"]:
                seen_convergence = True
                break
        assert seen_convergence and it >= 1
        # the whole trajectory: the reference's own scan2MapOptimization against the oracle's loop
        full_ref = r.optimize(case["pose_guess"]); full = o.optimize(case["pose_guess"])
        assert np.array_equal(full["pose"], full_ref["pose"]) and np.array_equal(full["pose"], tf)
        assert full["iterations"] == it + 1 and full["converged"] and full["degenerate"] == full_ref["degenerate"]


def test_degenerate_geometry_takes_the_projection_branch():
    """A map that constrains nothing along one axis: a single plane of surface points and no usable corners. The smallest eigenvalues of
    AtA fall below 100, isDegenerate is set, and the update is projected with matP = matV.inv() * matV2 (backMapping.cpp:1008-1036)."""
    g = np.random.default_rng(3)
    n = 6000
    xy = g.uniform(-20, 20, (n, 2))
    plane = np.concatenate([xy, np.zeros((n, 1)), np.zeros((n, 1))], axis=1).astype(np.float32)
    plane[:, 2] += g.normal(0, 0.002, n).astype(np.float32)
    corners = np.concatenate([g.uniform(-20, 20, (400, 2)), g.uniform(5, 6, (400, 1)), np.zeros((400, 1))], axis=1).astype(np.float32)
    scan_s = plane[g.choice(n, 1500, replace=False)].copy(); scan_s[:, 2] += np.float32(0.05)
    scan_c = corners[:60].copy(); scan_c[:, :3] += np.float32(30.0)        # nothing within 1 m: no corner row
    case = {"map_corner": corners, "map_surf": plane, "scan_corner": scan_c, "scan_surf": scan_s}
    o, r = _pair(case)
    tf = np.array([0.004, -0.003, 0.01, 0.05, -0.04, 0.02], np.float32)
    a = o.optimize(tf, edge_min=10, surf_min=100); b = r.optimize(tf, edge_min=10, surf_min=100)
    assert b["degenerate"] and a["degenerate"]
    assert np.array_equal(a["pose"], b["pose"])
    first = r.step(tf, 0); mine = o.optimize(tf, max_iter=1)
    assert first["degenerate"] and np.array_equal(first["pose"], mine["pose"])


def test_feature_count_gate_and_thread_invariance_of_the_oracle(small_case):
    o, r = _pair(small_case)
    few_o = o.optimize(small_case["pose_guess"], edge_min=10 ** 6); few_r = r.optimize(small_case["pose_guess"], edge_min=10 ** 6)
    assert few_o["iterations"] == -1 and not few_r["enough_features"]
    assert np.array_equal(few_o["pose"], small_case["pose_guess"]) and np.array_equal(few_r["pose"], small_case["pose_guess"])
    # the oracle is thread-count invariant (no shared bit-packed mask) and equal to the single-threaded reference
    o1 = ob.Scan2Map(threads=1); o1.setMap(small_case["map_corner"], small_case["map_surf"]); o1.setScan(small_case["scan_corner"], small_case["scan_surf"])
    ref = r.optimize(small_case["pose_guess"])["pose"]
    assert np.array_equal(o1.optimize(small_case["pose_guess"])["pose"], ref) and np.array_equal(o.optimize(small_case["pose_guess"])["pose"], ref)
