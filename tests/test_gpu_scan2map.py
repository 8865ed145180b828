"""GPU parity of the scan-to-map optimisation (backMapping.cpp:681-1054) against the CPU oracle, through the
C-ABI. Per-point work is float and restated operation by operation, so selection flags and coefficients must be BIT-EQUAL;
AtA/AtB are double sums in a different order rounded to float (<= 1 ulp apart); poses within 1e-4 rad / 1e-4 m."""
import numpy as np
import pytest

import helpers
from oracle import binding as ob
from rolo_b200 import synth
from rolo_b200.scan2map import Scan2Map

pytestmark = pytest.mark.gpu


def _both(case):
    o = ob.Scan2Map(); o.setMap(case["map_corner"], case["map_surf"]); o.setScan(case["scan_corner"], case["scan_surf"])
    g = Scan2Map(); g.setMap(case["map_corner"], case["map_surf"]); g.setScan(case["scan_corner"], case["scan_surf"])
    return o, g


@pytest.mark.parametrize("sensor,rings,cols", [("small", 32, 512), ("vlp64", 64, 1024)])
def test_one_evaluation_matches_the_oracle(sensor, rings, cols):
    case = synth.scan2map_case(helpers.oracle_features(rings, cols), sensor)
    o, g = _both(case)
    for tf in (case["pose_guess"], case["pose_true"], case["pose_guess"] + np.float32(0.02)):
        lo, lg = o.linearize(tf), g.linearize(tf)
        assert lo["n_sel"] == lg["n_sel"] > 1000
        for k in ("flag_corner", "flag_surf", "coeff_corner", "coeff_surf"):
            assert np.array_equal(lo[k], lg[k]), k
        scale = np.abs(lo["AtA"]).max()
        assert np.max(np.abs(lo["AtA"] - lg["AtA"])) <= 2e-7 * scale and np.allclose(lo["AtB"], lg["AtB"], rtol=0, atol=2e-7 * np.abs(lo["AtB"]).max())
        assert np.array_equal(lg["AtA"], lg["AtA"].T)


@pytest.mark.parametrize("sensor,rings,cols", [("small", 32, 512), ("vlp64", 64, 1024)])
def test_optimisation_matches_the_oracle(sensor, rings, cols):
    case = synth.scan2map_case(helpers.oracle_features(rings, cols), sensor)
    o, g = _both(case)
    ro, rg = o.optimize(case["pose_guess"]), g.optimize(case["pose_guess"])
    assert rg["enough_features"] and rg["converged"] == ro["converged"] and rg["degenerate"] == ro["degenerate"]
    assert rg["iterations"] == ro["iterations"]
    assert np.max(np.abs(rg["pose"][:3] - ro["pose"][:3])) <= 1e-4 and np.max(np.abs(rg["pose"][3:] - ro["pose"][3:])) <= 1e-4
    assert np.max(np.abs(rg["pose"][:3] - case["pose_true"][:3])) < 2e-3 and np.max(np.abs(rg["pose"][3:] - case["pose_true"][3:])) < 1e-2
    # same call again: deterministic (block-ordered reduction)
    assert np.array_equal(g.optimize(case["pose_guess"])["pose"], rg["pose"])


def test_edge_cases():
    case = synth.scan2map_case(helpers.oracle_features(32, 512), "small", n_keyframes=2)
    o, g = _both(case)
    # too few features: the reference warns and leaves the pose alone (backMapping.cpp:687,705-707)
    few = g.optimize(case["pose_guess"], edge_min=10 ** 6)
    assert not few["enough_features"] and few["iterations"] == 0 and np.array_equal(few["pose"], case["pose_guess"])
    # a scan far away from the map selects nothing; LMOptimization returns early 30 times (:956-958), pose unchanged
    far = case["pose_guess"].copy(); far[3] += 500.0
    lo, lg = o.linearize(far), g.linearize(far)
    assert lo["n_sel"] == lg["n_sel"] == 0 and not lg["flag_surf"].any()
    r = g.optimize(far)
    assert r["iterations"] == 30 and not r["converged"] and np.array_equal(r["pose"], far)
    assert o.optimize(far)["iterations"] == 30
    # empty / tiny maps (fewer than 5 points: nearestKSearch cannot return 5 neighbours)
    g2 = Scan2Map(); g2.setMap(case["map_corner"][:3], np.zeros((0, 4), np.float32)); g2.setScan(case["scan_corner"], case["scan_surf"])
    assert g2.linearize(case["pose_guess"])["n_sel"] == 0
    # degenerate geometry: a map made of one plane constrains 3 of 6 degrees of freedom -> isDegenerate, projected update
    flat = case["map_surf"][np.abs(case["map_surf"][:, 2] - np.median(case["map_surf"][:, 2])) < 0.3]
    sflat = case["scan_surf"]
    o3 = ob.Scan2Map(); o3.setMap(case["map_corner"][:5], flat); o3.setScan(case["scan_corner"], sflat)
    g3 = Scan2Map(); g3.setMap(case["map_corner"][:5], flat); g3.setScan(case["scan_corner"], sflat)
    ro, rg = o3.optimize(case["pose_guess"]), g3.optimize(case["pose_guess"])
    assert ro["degenerate"] and rg["degenerate"] and ro["iterations"] == rg["iterations"]
    assert np.max(np.abs(ro["pose"] - rg["pose"])) <= 1e-4


def test_full_size_frame_converges_to_the_truth():
    """128 x 2048 frames (BASELINE's size): size-independent property instead of an oracle run per point."""
    from rolo_b200.frontend import FrontEnd
    fe = FrontEnd(n_scan=128, horizon_scan=2048)
    def feats(raw):
        q = fe.feature_extraction(fe.image_projection(raw))
        return q["corner"], q["surface"]
    case = synth.scan2map_case(feats, "os1-128", n_keyframes=4)
    g = Scan2Map(); g.setMap(case["map_corner"], case["map_surf"]); g.setScan(case["scan_corner"], case["scan_surf"])
    r = g.optimize(case["pose_guess"])
    assert r["converged"] and np.max(np.abs(r["pose"][:3] - case["pose_true"][:3])) < 2e-3 and np.max(np.abs(r["pose"][3:] - case["pose_true"][3:])) < 1e-2
    o = ob.Scan2Map(); o.setMap(case["map_corner"], case["map_surf"]); o.setScan(case["scan_corner"], case["scan_surf"])
    ro = o.optimize(case["pose_guess"])
    assert np.max(np.abs(r["pose"] - ro["pose"])) <= 1e-4 and r["iterations"] == ro["iterations"]
