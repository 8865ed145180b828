"""Scan-to-map optimisation (backMapping.cpp:681-1054) without a GPU:
 * the oracle's restated third-party numerics against numpy / scipy,
 * the product's host algebra (s2m_math.cuh through the testing hooks) BIT-EQUAL to the oracle's,
 * the oracle's optimisation recovers a known pose on a synthetic map."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import capi, synth
import helpers

vp = lambda a: C.c_void_p(a.ctypes.data)


def _cases(seed, count):
    g = np.random.default_rng(seed)
    for _ in range(count):
        yield g


def test_oracle_numerics_against_numpy():
    L = ob.lib()
    g = np.random.default_rng(0)
    for n in (3, 6):
        for _ in range(40):
            B = g.standard_normal((n, n)).astype(np.float32); A = (B @ B.T).astype(np.float32)
            W = np.empty(n, np.float32); V = np.empty((n, n), np.float32)
            L.ro_s2m_eigen(n, vp(A), vp(W), vp(V))
            w = np.linalg.eigvalsh(A.astype(np.float64))[::-1]
            assert np.allclose(W, w, rtol=2e-5, atol=2e-5 * abs(w).max())
            assert np.allclose(V @ V.T, np.eye(n), atol=1e-5) and np.allclose(V @ A @ V.T, np.diag(W), atol=5e-5 * abs(w).max())
    for _ in range(40):
        A = g.standard_normal((5, 3)).astype(np.float32); b = -np.ones(5, np.float32); x = np.empty(3, np.float32)
        L.ro_s2m_qr_solve(5, 3, vp(A), vp(b), vp(x), 1)
        assert np.allclose(x, np.linalg.lstsq(A.astype(np.float64), b.astype(np.float64), rcond=None)[0], rtol=1e-4, atol=1e-4)
        A = g.standard_normal((6, 6)).astype(np.float32); A = A @ A.T + np.eye(6, dtype=np.float32); b = g.standard_normal(6).astype(np.float32)
        x = np.empty(6, np.float32); L.ro_s2m_qr_solve(6, 6, vp(A), vp(b), vp(x), 0)
        assert np.allclose(x, np.linalg.solve(A.astype(np.float64), b), rtol=2e-3, atol=2e-4)
        inv = np.empty((6, 6), np.float32)
        assert L.ro_s2m_invert(6, vp(A), vp(inv)) == 1 and np.allclose(inv, np.linalg.inv(A.astype(np.float64)), rtol=2e-3, atol=2e-4)
    from scipy.spatial.transform import Rotation
    tf = np.array([0.1, -0.2, 0.7, 1, 2, 3], np.float32); t = np.empty(12, np.float32)
    L.ro_s2m_get_transformation(vp(tf), vp(t))
    assert np.allclose(t.reshape(3, 4)[:, :3], Rotation.from_euler("ZYX", [0.7, -0.2, 0.1]).as_matrix(), atol=1e-6)
    assert np.array_equal(t.reshape(3, 4)[:, 3], tf[3:])


def test_product_host_algebra_is_bit_equal_to_the_oracle():
    L = ob.lib(); P = capi.load()
    g = np.random.default_rng(5)
    for n in (3, 6):
        for _ in range(60):
            B = g.standard_normal((n, n)).astype(np.float32) * g.uniform(0.01, 30); A = (B @ B.T).astype(np.float32)
            Wo = np.empty(n, np.float32); Vo = np.empty((n, n), np.float32); Wp = np.empty_like(Wo); Vp = np.empty_like(Vo)
            L.ro_s2m_eigen(n, vp(A), vp(Wo), vp(Vo)); P.rgt_s2m_eigen(n, vp(A), vp(Wp), vp(Vp))
            assert np.array_equal(Wo, Wp) and np.array_equal(Vo, Vp)
    for _ in range(60):
        A = (g.standard_normal((5, 3)) * g.uniform(0.1, 50) + g.uniform(-20, 20, 3)).astype(np.float32); b = -np.ones(5, np.float32)
        xo = np.empty(3, np.float32); xp = np.empty(3, np.float32)
        L.ro_s2m_qr_solve(5, 3, vp(A), vp(b), vp(xo), 1); P.rgt_s2m_qr_solve(5, 3, vp(A), vp(b), vp(xp))
        assert np.array_equal(xo, xp)
        A = g.standard_normal((6, 6)).astype(np.float32); A = (A @ A.T * 100).astype(np.float32); b = g.standard_normal(6).astype(np.float32)
        xo = np.empty(6, np.float32); xp = np.empty(6, np.float32)
        L.ro_s2m_qr_solve(6, 6, vp(A), vp(b), vp(xo), 0); P.rgt_s2m_qr_solve(6, 6, vp(A), vp(b), vp(xp))
        assert np.array_equal(xo, xp)
        io = np.empty((6, 6), np.float32); ip = np.empty((6, 6), np.float32)
        assert L.ro_s2m_invert(6, vp(A), vp(io)) == P.rgt_s2m_invert6(vp(A), vp(ip)) == 1 and np.array_equal(io, ip)
        tf = g.uniform(-3, 3, 6).astype(np.float32); to = np.empty(12, np.float32); tp = np.empty(12, np.float32)
        L.ro_s2m_get_transformation(vp(tf), vp(to)); P.rgt_s2m_get_transformation(vp(tf), vp(tp))
        assert np.array_equal(to, tp)
    sing = np.zeros((6, 6), np.float32); io = np.empty((6, 6), np.float32)
    assert L.ro_s2m_invert(6, vp(sing), vp(io)) == P.rgt_s2m_invert6(vp(sing), vp(io)) == 0


def test_oracle_recovers_the_pose_on_a_synthetic_map():
    case = synth.scan2map_case(helpers.oracle_features(32, 512), "small")
    o = ob.Scan2Map(); o.setMap(case["map_corner"], case["map_surf"]); o.setScan(case["scan_corner"], case["scan_surf"])
    lin = o.linearize(case["pose_guess"])
    assert lin["n_sel"] == int(lin["flag_corner"].sum()) + int(lin["flag_surf"].sum()) > 1000
    assert np.allclose(lin["AtA"], lin["AtA"].T)
    res = o.optimize(case["pose_guess"])
    assert res["converged"] and not res["degenerate"] and 1 <= res["iterations"] <= 30
    assert np.max(np.abs(res["pose"][:3] - case["pose_true"][:3])) < 2e-3 and np.max(np.abs(res["pose"][3:] - case["pose_true"][3:])) < 1e-2
    # thread-count invariance (the per-point work is independent; sums run in index order)
    o1 = ob.Scan2Map(threads=1); o1.setMap(case["map_corner"], case["map_surf"]); o1.setScan(case["scan_corner"], case["scan_surf"])
    assert np.array_equal(o1.optimize(case["pose_guess"])["pose"], res["pose"])
    # too few features: the reference only warns and leaves the pose alone (backMapping.cpp:687,705-707)
    few = o.optimize(case["pose_guess"], edge_min=10 ** 6)
    assert few["iterations"] == -1 and np.array_equal(few["pose"], case["pose_guess"])


def _rot(tf):
    return synth.pose6_matrix(np.asarray(tf, dtype=np.float64))


def test_coefficients_and_jacobian_have_their_geometric_meaning():
    """Independent (numpy, float64) reading of what the restated formulas must mean: a surface coefficient is s x (unit normal of
    the plane through the five neighbours, signed distance); a corner coefficient is s x (unit vector from the line to the point,
    distance to the line); the rows of matA are the derivatives of the residual with respect to (roll, pitch, yaw, x, y, z)."""
    case = synth.scan2map_case(helpers.oracle_features(32, 512), "small")
    o = ob.Scan2Map(); o.setMap(case["map_corner"], case["map_surf"]); o.setScan(case["scan_corner"], case["scan_surf"])
    tf = case["pose_guess"].astype(np.float64)
    lin = o.linearize(case["pose_guess"])
    T = _rot(tf)
    # --- surfaces
    P = case["scan_surf"][:, :3].astype(np.float64); W = P @ T[:3, :3].T + T[:3, 3]
    M = case["map_surf"][:, :3].astype(np.float64)
    sel = np.flatnonzero(lin["flag_surf"])[:300]
    for i in sel:
        d = np.sum((M - W[i]) ** 2, axis=1); nn = np.argsort(d, kind="stable")[:5]
        n = np.linalg.lstsq(M[nn], -np.ones(5), rcond=None)[0]; nrm = np.linalg.norm(n); n, dd = n / nrm, 1.0 / nrm
        pd2 = n @ W[i] + dd
        s = 1 - 0.9 * abs(pd2) / np.sqrt(np.linalg.norm(P[i]))
        c = lin["coeff_surf"][i].astype(np.float64)
        assert np.allclose(c[:3], s * n, atol=2e-3) and abs(c[3] - s * pd2) < 2e-3
    # --- corners
    P = case["scan_corner"][:, :3].astype(np.float64); Wc = P @ T[:3, :3].T + T[:3, 3]
    Mc = case["map_corner"][:, :3].astype(np.float64)
    selc = np.flatnonzero(lin["flag_corner"])[:100]
    assert selc.size > 20
    for i in selc:
        d = np.sum((Mc - Wc[i]) ** 2, axis=1); nn = np.argsort(d, kind="stable")[:5]
        ctr = Mc[nn].mean(axis=0); w, v = np.linalg.eigh(np.cov((Mc[nn] - ctr).T, bias=True)); u = v[:, 2]
        r = Wc[i] - ctr; perp = r - (r @ u) * u; dist = np.linalg.norm(perp)
        s = 1 - 0.9 * dist
        c = lin["coeff_corner"][i].astype(np.float64)
        assert abs(c[3] - s * dist) < 2e-3 and np.allclose(c[:3], s * perp / dist, atol=5e-3)
    # --- Jacobian rows by central differences of r(tf) = c . (R(tf) p + t(tf)) -> AtA, AtB
    rows = []; res = []
    for pts, cf, fl in ((case["scan_corner"], lin["coeff_corner"], lin["flag_corner"]), (case["scan_surf"], lin["coeff_surf"], lin["flag_surf"])):
        idx = np.flatnonzero(fl); p = pts[idx, :3].astype(np.float64); c = cf[idx].astype(np.float64)
        J = np.zeros((idx.size, 6))
        for k in range(6):
            h = 1e-6; tp = tf.copy(); tm = tf.copy(); tp[k] += h; tm[k] -= h
            Tp, Tm = _rot(tp), _rot(tm)
            J[:, k] = np.sum(c[:, :3] * ((p @ Tp[:3, :3].T + Tp[:3, 3]) - (p @ Tm[:3, :3].T + Tm[:3, 3])), axis=1) / (2 * h)
        rows.append(J); res.append(-c[:, 3])
    J = np.vstack(rows); b = np.concatenate(res)
    AtA, AtB = J.T @ J, J.T @ b
    assert np.allclose(lin["AtA"], AtA, rtol=2e-3, atol=2e-3 * np.abs(AtA).max())
    assert np.allclose(lin["AtB"], AtB, rtol=2e-3, atol=2e-3 * np.abs(AtB).max())


def test_restated_opencv_routines_against_real_opencv_golden_vectors():
    """tests/golden/opencv_s2m_golden.npz holds outputs of the real OpenCV (tools/make_opencv_golden.py, the image's cv2 4.13):
    the restated cv::eigen (3x3, 6x6) and cv::Mat::inv are BIT-EQUAL to it; cv::solve(DECOMP_QR) dispatches to LAPACK in that
    build, so the Householder restatement agrees to rounding (1e-5 relative on these well-conditioned systems)."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "opencv_s2m_golden.npz"))
    L = ob.lib(); P = capi.load()
    for n in (3, 6):
        for A, W, V in zip(G[f"eig{n}_A"], G[f"eig{n}_W"], G[f"eig{n}_V"]):
            A = np.ascontiguousarray(A)
            for fn in (L.ro_s2m_eigen, P.rgt_s2m_eigen):                      # oracle and product host code
                w = np.empty(n, np.float32); v = np.empty((n, n), np.float32)
                fn(n, vp(A), vp(w), vp(v))
                assert np.array_equal(w, W) and np.array_equal(v, V)
    for M, I in zip(G["inv_M"], G["inv_I"]):
        M = np.ascontiguousarray(M); inv = np.empty((6, 6), np.float32)
        assert L.ro_s2m_invert(6, vp(M), vp(inv)) == 1 and np.array_equal(inv, I)
        assert P.rgt_s2m_invert6(vp(M), vp(inv)) == 1 and np.array_equal(inv, I)
    for A, b, x in zip(G["qr_A"], G["qr_b"], G["qr_x"]):
        A = np.ascontiguousarray(A); b = np.ascontiguousarray(b); xo = np.empty(6, np.float32)
        L.ro_s2m_qr_solve(6, 6, vp(A), vp(b), vp(xo), 0)
        assert np.max(np.abs(xo - x)) <= 1e-5 * np.max(np.abs(x))


def test_restated_opencv_routines_against_live_opencv():
    cv2 = pytest.importorskip("cv2")
    L = ob.lib()
    g = np.random.default_rng(77)
    for _ in range(100):
        P5 = (g.standard_normal((5, 3)) * g.uniform(0.05, 2.0)).astype(np.float32); P5 -= P5.mean(0)
        A = (P5.T @ P5 / 5).astype(np.float32); A = ((A + A.T) / 2).astype(np.float32)
        w = np.empty(3, np.float32); v = np.empty((3, 3), np.float32); L.ro_s2m_eigen(3, vp(A), vp(w), vp(v))
        ok, wc, vc = cv2.eigen(A)
        assert np.array_equal(w, wc.reshape(-1)) and np.array_equal(v, vc)
