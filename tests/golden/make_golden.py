"""Generates tests/golden/*.npz with the CPU oracle (oracle/) on seeded synthetic inputs.

The reference ships no golden vectors for this path and cannot be built in this image, so these
fixtures are produced by the oracle restatement itself: they pin the oracle (and through it the CUDA path) against
regressions, and they travel to the GPU box where /root/reference does not exist.

    python tests/golden/make_golden.py        # rewrites the fixtures (deterministic: single-threaded oracle)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import binding as ob            # noqa: E402
from rolo_b200 import synth                 # noqa: E402


def registration_fixture():
    tgt, src, T_gt = synth.scan_pair("tiny")            # 16 x 256 = 4096 points per cloud, seeds 1234 / 5678
    out = {"T_gt": T_gt, "target_sha": np.frombuffer(tgt.tobytes(), dtype=np.uint8)[::997].copy(),
           "source_sha": np.frombuffer(src.tobytes(), dtype=np.uint8)[::997].copy()}
    for name, opt, iters, eps in (("gn", ob.OPT_GAUSS_NEWTON, 6, 1e-30), ("lm", ob.OPT_LM, None, None), ("so3lm", ob.OPT_SO3_LM, None, None)):
        o = ob.Oracle(); o.setNumThreads(1)
        o.setPolarResolution(0.175, 0.175, 2.0); o.setOptimizerType(opt)
        if iters:
            o.setMaximumIterations(iters); o.setRotationEpsilon(eps); o.setTransformationEpsilon(eps)
        o.setInputTarget(tgt); o.setInputSource(src)
        T = o.align()
        out[f"{name}_final"] = o.getFinalPoseDouble(); out[f"{name}_final_f32"] = T
        out[f"{name}_trace"] = o.getTrace(); out[f"{name}_iters"] = np.array(o.nrIterations()); out[f"{name}_converged"] = np.array(o.hasConverged())
        if name == "so3lm":
            t = o.computeTranslation(np.zeros(3), T[:3, 3].astype(np.float64), np.array([0.01, 0.0, 0.0]), 0.1, 0.1, 0.3)
            out["so3lm_translation"] = t; out["so3lm_translation_trace"] = o.getTrace()
    # stage-level vectors
    o = ob.Oracle(); o.setNumThreads(1); o.setPolarResolution(0.175, 0.175, 2.0); o.setInputTarget(tgt); o.setInputSource(src)
    o.calculateCovariances(0); o.calculateCovariances(1)
    out["knn_target_first64"] = o.knn(1, 20)[:64]
    out["cov_source_first64"] = o.getSourceCovariances()[:64]
    o.createVoxelmap()
    k, m, c, n = o.getVoxels()
    out["vox_keys"] = k; out["vox_mean"] = m; out["vox_npts"] = n; out["vox_cov_trace"] = np.trace(c[:, :3, :3], axis1=1, axis2=2)
    for method, tag in ((ob.DIRECT1, "d1"), (ob.DIRECT7, "d7")):
        o.setNeighborSearchMethod(method)
        cst, H, b = o.linearize(T_gt)
        out[f"lin_{tag}_cost"] = np.array(cst); out[f"lin_{tag}_H"] = H; out[f"lin_{tag}_b"] = b
    return out


def frontend_fixture():
    scene = synth.make_scene()
    raw = synth.raw_frame(scene, synth.pose((1.0, -2.0, 30.0), (1.0, 0.5, 0.0)), "small", extra=300)
    p = ob.image_projection(raw, 32, 512)
    f = ob.feature_extraction(p, 32, 512)
    return {"n": np.array(p["n"]), "start_ring_index": p["start_ring_index"], "end_ring_index": p["end_ring_index"],
            "point_col_ind_head": p["point_col_ind"][:256], "point_range_head": p["point_range"][:256],
            "label_hist": np.bincount(f["label"] + 1, minlength=3), "corner": f["corner"], "surface_head": f["surface"][:128],
            "n_surface": np.array(len(f["surface"])), "curvature_head": f["curvature"][:256]}


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "registration_tiny.npz"), **registration_fixture())
    np.savez_compressed(os.path.join(HERE, "frontend_small.npz"), **frontend_fixture())
    for f in ("registration_tiny.npz", "frontend_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
