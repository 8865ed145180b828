"""One full-size (128x2048) registration, for ncu launch lists / ncu --set full captures.
usage: python tools/profile_align.py [pairs] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP, RotVGICPBatch

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sensor = os.environ.get("ROLO_SENSOR", "os1-128")
scene = synth.make_scene()
rng = np.random.default_rng(99)
Ts = [synth.random_offset(rng) for _ in range(pairs)]
clouds = [(synth.scan(scene, np.eye(4), sensor, seed=1234 + i), synth.scan(scene, Ts[i], sensor, seed=5678 + i)) for i in range(pairs)]
b = RotVGICPBatch(pairs)
for i in range(pairs):
    r = b[i]
    r.setPolarResolution(0.175, 0.175, 2.0); r.setOptimizerType(capi.OPT_GAUSS_NEWTON); r.setMaximumIterations(20)
    r.setRotationEpsilon(1e-30); r.setTransformationEpsilon(1e-30)
for rep in range(reps):
    t0 = time.perf_counter()
    for i in range(pairs):
        b[i].setInputTarget(clouds[i][0]); b[i].setInputSource(clouds[i][1])
    t1 = time.perf_counter()
    out = b.align()
    t2 = time.perf_counter()
    print(f"rep {rep}: upload {1e3*(t1-t0):.2f} ms, align {1e3*(t2-t1):.2f} ms ({pairs} pairs), linearize {b.lastLinearizeUs():.1f} us x {b.lastLinearizeLaunches()}")
err = [synth.pose_error(out[i].astype(np.float64), Ts[i]) for i in range(pairs)]
print("max err vs gt", max(e[0] for e in err), max(e[1] for e in err))
