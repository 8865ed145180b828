"""Development probe: where does the end-to-end (pinned host inputs) step spend its time compared with the resident one?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICPBatch
import bench
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
scene = synth.make_scene(); Ts = bench.offsets_for(0, P)
td, sd, th, sh = [], [], [], []
for i in range(P):
    t = synth.scan_torch(scene, np.eye(4), "os1-128", seed=1234 + i, device=dev); s = synth.scan_torch(scene, Ts[i], "os1-128", seed=5678 + i, device=dev)
    td.append(t); sd.append(s); th.append(t.cpu().pin_memory()); sh.append(s.cpu().pin_memory())
b = RotVGICPBatch(P, device=0)
for i in range(P): bench.configure(b[i], capi)
for mode in ("resident", "pinned", "resident", "pinned", "pinned"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(P):
        if mode == "resident": b[i].setInputTarget(td[i]); b[i].setInputSource(sd[i])
        else: b[i].setInputTarget(th[i]); b[i].setInputSource(sh[i])
    t1 = time.perf_counter(); b.align(); t2 = time.perf_counter()
    st = b.lastStages()
    print(f"{mode:9s} install {1e3*(t1-t0):6.2f} ms  align {1e3*(t2-t1):7.2f} ms  | cov {st['covariances_ms']:.1f} knn {st['knn_kernels_ms']:.1f} vox {st['voxel_maps_ms']:.1f} iter {st['iterations_ms']:.1f}")
