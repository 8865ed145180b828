"""Writes tests/golden/opencv_s2m_golden.npz: inputs and outputs of the REAL OpenCV (the image's cv2) for the routines the
scan-to-map path calls and the oracle restates — cv::eigen on symmetric CV_32F 3x3 / 6x6, cv::Mat::inv (DECOMP_LU) 6x6,
cv::solve(DECOMP_QR) 6x6 — so the restatement stays pinned where cv2 is absent. Deterministic (seeded)."""
import os
import numpy as np
import cv2

g = np.random.default_rng(20240924)
out = {"cv2_version": np.array(cv2.__version__)}
for n in (3, 6):
    As, Ws, Vs = [], [], []
    for t in range(60):
        if n == 3:      # covariance of five points, like cornerOptimization
            P = (g.standard_normal((5, 3)) * g.uniform(0.05, 2.0)).astype(np.float32); P -= P.mean(0)
            A = (P.T @ P / 5).astype(np.float32)
        else:           # normal-equation-like
            J = (g.standard_normal((400, 6)) * g.uniform(0.1, 3.0, 6)).astype(np.float32)
            A = (J.T @ J).astype(np.float32)
        A = ((A + A.T) / 2).astype(np.float32)
        ok, w, v = cv2.eigen(A)
        As.append(A); Ws.append(w.reshape(-1)); Vs.append(v)
    out[f"eig{n}_A"], out[f"eig{n}_W"], out[f"eig{n}_V"] = np.array(As), np.array(Ws), np.array(Vs)
Ms, Is, Ss, Bs, Xs = [], [], [], [], []
for t in range(60):
    M = g.standard_normal((6, 6)).astype(np.float32)
    ok, inv = cv2.invert(M)
    Ms.append(M); Is.append(inv)
    J = (g.standard_normal((400, 6)) * g.uniform(0.3, 2.0, 6)).astype(np.float32)
    S = (J.T @ J).astype(np.float32); b = g.standard_normal((6, 1)).astype(np.float32) * 10
    ok, x = cv2.solve(S, b, flags=cv2.DECOMP_QR)
    Ss.append(S); Bs.append(b.reshape(-1)); Xs.append(x.reshape(-1))
out["inv_M"], out["inv_I"], out["qr_A"], out["qr_b"], out["qr_x"] = map(np.array, (Ms, Is, Ss, Bs, Xs))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "opencv_s2m_golden.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
