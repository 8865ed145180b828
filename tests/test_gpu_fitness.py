"""Loop-closure verification: exact nearest-neighbour queries of external points and
pcl::Registration::getFitnessScore against the oracle's kd-tree — indices and float distances bit-equal, score equal."""
import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP

pytestmark = pytest.mark.gpu


def _pair(sensor, seed=3):
    scene = synth.make_scene()
    T = synth.pose((1.0, -2.0, 4.0), (0.3, -0.2, 0.05))
    return synth.scan(scene, np.eye(4), sensor, seed=seed), synth.scan(scene, T, sensor, seed=seed + 1), T


@pytest.mark.parametrize("sensor", ["small", "vlp64"])
def test_nearest_and_fitness_match_the_oracle(sensor):
    tgt, src, T = _pair(sensor)
    o = ob.Oracle(); o.setInputTarget(tgt); o.setInputSource(src)
    g = RotVGICP(); g.setInputTarget(tgt); g.setInputSource(src)
    far = synth.pose((0, 0, 0), (300.0, -40.0, 10.0))      # most queries leave the target's bounding box
    for M in (np.eye(4), T, np.linalg.inv(T), far):
        M32 = M.astype(np.float32)
        so, io, do = o.getFitnessScore(M32, n_source=src.shape[0])
        ig, dg = g.nearestTarget(M32, src.shape[0])
        assert np.array_equal(do, dg)
        assert np.array_equal(io, ig)
        assert g.getFitnessScore(M32) == so
        for mr in (0.01, 0.25):
            assert g.getFitnessScore(M32, mr) == o.getFitnessScore(M32, mr)
    assert g.getFitnessScore(T.astype(np.float32)) < g.getFitnessScore(np.eye(4, dtype=np.float32))
    assert g.getFitnessScore(np.eye(4, dtype=np.float32), 0.0) == np.finfo(np.float64).max     # nothing within range


def test_loop_closure_flow_with_the_gicp_core():
    """backMapping.cpp:2331-2352 with the GICP core in 6-DoF LM mode in place of pcl::IterativeClosestPoint: align, then accept on
    convergence and fitness (historyKeyframeFitnessScore = 0.3 in config/params.yaml)."""
    tgt, src, T = _pair("vlp64", seed=11)
    g = RotVGICP(); g.setPolarResolution(0.175, 0.175, 2.0); g.setOptimizerType(capi.OPT_LM); g.setMaximumIterations(100)
    g.setInputTarget(tgt); g.setInputSource(src)
    before = g.getFitnessScore(np.eye(4, dtype=np.float32))
    Tf = g.align()
    rot, tr = synth.pose_error(Tf.astype(np.float64), T)
    assert g.hasConverged() and rot < 2e-3 and tr < 2e-2
    after = g.getFitnessScore()
    assert after < before and after <= 1.05 * g.getFitnessScore(T.astype(np.float32))
    o = ob.Oracle(); o.setInputTarget(tgt); o.setInputSource(src)
    assert after == o.getFitnessScore(Tf)
