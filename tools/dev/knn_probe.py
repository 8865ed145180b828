"""Development probe: covariance estimation of one 128x2048 pair with the chunk kernel's work counters (ROLO_KNN_STATS=1)."""
import os, sys, time
os.environ.setdefault("ROLO_KNN_STATS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP

sensor = sys.argv[1] if len(sys.argv) > 1 else "os1-128"
scene = synth.make_scene()
dev = torch.device("cuda", 0)
t = synth.scan_torch(scene, np.eye(4), sensor, seed=1234, device=dev)
s = synth.scan_torch(scene, synth.T_GT_DEFAULT, sensor, seed=5678, device=dev)
g = RotVGICP(device=0)
g.setPolarResolution(0.175, 0.175, 2.0)
for it in range(3):
    g.setInputTarget(t); g.setInputSource(s)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g.calculate_covariances(capi.TARGET)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    g.calculate_covariances(capi.SOURCE)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"target {1e3*(t1-t0):.3f} ms  source {1e3*(t2-t1):.3f} ms (wall, incl. stats read-back)")
