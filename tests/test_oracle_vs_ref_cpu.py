"""CPU: the oracle restatement against oracle/_ref — the REFERENCE's own rot_gicp sources compiled from
/root/reference/include (nothing copied) with oracle/ref_shim standing in for Eigen/PCL/Boost (absent from the image).
This is what pins the restatement of rot_vgicp_impl.hpp / lsq_registration_impl.hpp / vmp_voxel.hpp / so3.hpp line by
line. Where /root/reference exists the library is built on demand and its absence is an error; only a machine without the checkout
and without the prebuilt file skips."""
import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import synth

ob.require_ref(ob.ref_available, "oracle/_ref/librolo_ref.so")


def _pair(cls, tgt, src, uniform=None, method=ob.DIRECT1, reg=ob.REG_PLANE):
    r = cls()
    r.setNumThreads(4)
    if uniform:
        r.setResolution(uniform)
    else:
        r.setPolarResolution(0.175, 0.175, 2.0)
    r.setNeighborSearchMethod(method); r.setRegularizationMethod(reg)
    r.setInputTarget(tgt); r.setInputSource(src)
    return r


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("reg", [ob.REG_PLANE, ob.REG_NONE, ob.REG_MIN_EIG, ob.REG_NORMALIZED_MIN_EIG, ob.REG_FROBENIUS, ob.REG_PLANE_S])
def test_covariances(tiny_pair, reg):
    tgt, src, _ = tiny_pair
    o = _pair(ob.Oracle, tgt, src, reg=reg); r = _pair(ob.Reference, tgt, src, reg=reg)
    assert o.calculateCovariances(0) == 0 and r.calculateCovariances(0) == 0
    a, b = o.getSourceCovariances(), r.getSourceCovariances()
    assert a.shape == b.shape and np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b).max(axis=(1, 2), keepdims=True))) <= 1e-12


@pytest.mark.parametrize("grid,method", [("polar", ob.DIRECT1), ("polar", ob.DIRECT7), ("uniform", ob.DIRECT1), ("uniform", ob.DIRECT27)])
def test_per_call_quantities(tiny_pair, grid, method):
    tgt, src, T_gt = tiny_pair
    uni = 1.0 if grid == "uniform" else None
    o = _pair(ob.Oracle, tgt, src, uni, method); r = _pair(ob.Reference, tgt, src, uni, method)
    for x in (o, r):
        x.calculateCovariances(0); x.calculateCovariances(1)
    for T in (np.eye(4), T_gt, synth.pose((1.0, -2.0, 4.0), (0.15, -0.05, 0.02))):
        co, Ho, bo = o.linearize(T); cr, Hr, br = r.linearize(T)
        assert rel(co, cr) <= 1e-11 and rel(Ho, Hr) <= 1e-11 and rel(bo, br) <= 1e-10
        assert o.L.ro_num_correspondences(o.h) == r.numCorrespondences()
        c3, H3, b3 = o.so3_linearize(T); c3r, H3r, b3r = r.so3_linearize(T)
        assert rel(c3, c3r) <= 1e-11 and rel(H3, H3r) <= 1e-11 and rel(b3, b3r) <= 1e-10
        Tt = synth.pose((0.2, -0.1, 0.3), (0.01, 0.02, -0.01)) @ T
        assert rel(o.compute_error(Tt), r.compute_error(Tt)) <= 1e-11
    # translation stage incl. the Q3 behaviour of `last_transform.matrix().col(3).head<3>() = last_t0`
    o.setCtLambda(0.3); r.setCtLambda(0.3)
    g = np.array([0.2, -0.1, 0.05]); l = np.array([0.18, -0.12, 0.04])
    for t in (np.zeros(3), np.array([0.01, -0.02, 0.005])):
        co, Ho, bo = o.t3_linearize(t, g, l, 0.1, 0.1); cr, Hr, br = r.t3_linearize(t, g, l, 0.1, 0.1)
        assert rel(co, cr) <= 1e-11 and rel(Ho, Hr) <= 1e-11 and rel(bo, br) <= 1e-10
        assert rel(o.compute_t_error(t, g, l, 0.1, 0.1), r.compute_t_error(t, g, l, 0.1, 0.1)) <= 1e-11


@pytest.mark.parametrize("opt,iters,eps", [(ob.OPT_GAUSS_NEWTON, 8, 1e-30), (ob.OPT_LM, None, None), (ob.OPT_SO3_LM, None, None)])
def test_align_and_translation(small_pair, opt, iters, eps):
    tgt, src, _ = small_pair
    o = _pair(ob.Oracle, tgt, src); r = _pair(ob.Reference, tgt, src)
    for x in (o, r):
        x.setOptimizerType(opt)
        if iters:
            x.setMaximumIterations(iters); x.setRotationEpsilon(eps); x.setTransformationEpsilon(eps)
    To = o.align(); Tr = r.align()
    assert np.max(np.abs(To - Tr)) <= 1e-6            # float32 final_transformation_
    assert o.nrIterations() == r.nrIterations() and o.hasConverged() == r.hasConverged()
    if opt != ob.OPT_SO3_LM:
        assert rel(o.getFinalHessian(), r.getFinalHessian()) <= 1e-8
    to = o.computeTranslation(np.zeros(3), To[:3, 3].astype(np.float64), np.array([0.01, 0.0, 0.0]), 0.1, 0.1, 0.3)
    tr = r.computeTranslation(np.zeros(3), Tr[:3, 3].astype(np.float64), np.array([0.01, 0.0, 0.0]), 0.1, 0.1, 0.3)
    assert np.max(np.abs(to - tr)) <= 1e-9
