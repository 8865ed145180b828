"""CPU: the oracle reproduces the committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np

from oracle import binding as ob
from rolo_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_inputs_are_the_seeded_ones():
    g = np.load(os.path.join(G, "registration_tiny.npz"))
    tgt, src, T_gt = synth.scan_pair("tiny")
    assert np.array_equal(np.frombuffer(tgt.tobytes(), dtype=np.uint8)[::997], g["target_sha"])
    assert np.array_equal(np.frombuffer(src.tobytes(), dtype=np.uint8)[::997], g["source_sha"])
    assert np.allclose(T_gt, g["T_gt"])


def test_oracle_reproduces_registration_golden():
    g = np.load(os.path.join(G, "registration_tiny.npz"))
    tgt, src, T_gt = synth.scan_pair("tiny")
    for name, opt, iters, eps in (("gn", ob.OPT_GAUSS_NEWTON, 6, 1e-30), ("lm", ob.OPT_LM, None, None), ("so3lm", ob.OPT_SO3_LM, None, None)):
        o = ob.Oracle(); o.setNumThreads(4)            # thread count changes summation order only
        o.setPolarResolution(0.175, 0.175, 2.0); o.setOptimizerType(opt)
        if iters:
            o.setMaximumIterations(iters); o.setRotationEpsilon(eps); o.setTransformationEpsilon(eps)
        o.setInputTarget(tgt); o.setInputSource(src)
        o.align()
        assert np.allclose(o.getFinalPoseDouble(), g[f"{name}_final"], rtol=0, atol=1e-10)
        tr = o.getTrace()
        assert tr.shape == g[f"{name}_trace"].shape and np.array_equal(tr[:, 0], g[f"{name}_trace"][:, 0])
        assert np.allclose(tr[:, 1], g[f"{name}_trace"][:, 1], rtol=1e-9)
        assert o.nrIterations() == int(g[f"{name}_iters"]) and o.hasConverged() == bool(g[f"{name}_converged"])


def test_oracle_reproduces_frontend_golden():
    g = np.load(os.path.join(G, "frontend_small.npz"))
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, synth.pose((1.0, -2.0, 30.0), (1.0, 0.5, 0.0)), "small", extra=300)
    p = ob.image_projection(raw, 32, 512)
    f = ob.feature_extraction(p, 32, 512)
    assert p["n"] == int(g["n"]) and np.array_equal(p["start_ring_index"], g["start_ring_index"])
    assert np.array_equal(p["point_col_ind"][:256], g["point_col_ind_head"]) and np.array_equal(p["point_range"][:256], g["point_range_head"])
    assert np.array_equal(np.bincount(f["label"] + 1, minlength=3), g["label_hist"])
    assert np.array_equal(f["corner"], g["corner"]) and np.array_equal(f["surface"][:128], g["surface_head"])
