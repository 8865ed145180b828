"""Development: touches every kernel of the library once (for an ncu pass): batch of 4 pairs, front-end frame, scan-to-map, Scan Context, fitness."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP, RotVGICPBatch
import bench
dev = torch.device("cuda", 0)
scene = synth.make_scene(); P = 4
Ts = bench.offsets_for(0, P)
b = RotVGICPBatch(P, device=0)
for i in range(P):
    bench.configure(b[i], capi)
    b[i].setInputTarget(synth.scan_torch(scene, np.eye(4), "os1-128", seed=1234 + i, device=dev))
    b[i].setInputSource(synth.scan_torch(scene, Ts[i], "os1-128", seed=5678 + i, device=dev))
for _ in range(2):
    for i in range(P):
        b[i].setInputTarget(b[i]._keep[capi.TARGET]); b[i].setInputSource(b[i]._keep[capi.SOURCE])
    b.align()
r = RotVGICP(); r.setPolarResolution(0.175, 0.175, 2.0)
raw0 = synth.raw_frame(scene, synth.pose((0, 0, 0), (0, 0, 0)), "os1-128", seed=100)
raw1 = synth.raw_frame(scene, synth.pose((0, 0, 0.2), (0.05, 0.01, 0)), "os1-128", seed=101)
r.frontEndFrame(raw0, capi.TARGET, n_scan=128, horizon_scan=2048); r.swapSourceAndTarget(); r.frontEndFrame(raw1, capi.TARGET, n_scan=128, horizon_scan=2048)
T = r.align(); r.getFitnessScore(T)
print("done")
