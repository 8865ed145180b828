"""Development probe: covariance estimation of a merged multi-scan submap with the chunk kernel's work counters (ROLO_KNN_STATS=1)."""
import os, sys, time
os.environ.setdefault("ROLO_KNN_STATS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from rolo_b200 import capi, synth
from rolo_b200.registration import RotVGICP

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 77
scene = synth.make_scene(); dev = torch.device("cuda", 0)
clouds = []
for k in range(n_scans):
    a = 2 * np.pi * k / 77
    T = synth.pose((0.0, 0.0, np.degrees(a) + 90.0), (5.0 + 18.0 * np.cos(a), 3.0 + 10.0 * np.sin(a), 0.0))
    p = synth.scan_torch(scene, T, "os1-128x1024", sigma=0.004, seed=1000 + k, device=dev)
    Tm = torch.as_tensor(T, dtype=torch.float32, device=dev)
    q = p.clone(); q[:, :3] = p[:, :3] @ Tm[:3, :3].T + Tm[:3, 3]
    clouds.append(q)
target = torch.cat(clouds, dim=0).contiguous()
g = RotVGICP(device=0); g.setResolution(1.0)
for it in range(2):
    g.setInputTarget(target)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g.calculate_covariances(capi.TARGET)
    torch.cuda.synchronize(); print(f"target {target.shape[0]} pts: {1e3*(time.perf_counter()-t0):.2f} ms")
