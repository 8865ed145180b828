"""Writes tests/golden/cloud_info_digest.json: size + sha256 of two oracle-serialised rolo/CloudInfoStamp messages
(inputs are the seeded products of tests/test_cloud_info_cpu.py::fake_products)."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import binding as ob
from test_cloud_info_cpu import fake_products

out = {}
for stage in (0, 1):
    proj, feat = fake_products(seed=21 + stage, n_scan=8, n=301, nc=40, ns=77)
    msg = ob.cloud_info_serialize(stage, 5, 100, 200, "base_link", proj, feat, np.arange(9, dtype=np.float32), np.ones(36, np.float32), stage == 1)
    out[f"stage{stage}"] = {"size": len(msg), "sha256": hashlib.sha256(msg).hexdigest()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "cloud_info_digest.json"), "w"), indent=1)
print(out)
