"""CPU (no GPU): the product's HOST logic — the resumable GN/LM/SO3-LM/translation-LM state machine of
rolo_b200/csrc/lsq_host.hpp and the persistent-CTA schedule of linearize.cu — exercised through the test hooks of
include/rolo_gicp_testing.h. The per-evaluation numbers come from the oracle, so the driver must retrace the oracle's own
align()/computeTranslation() exactly (same requests, same accept/reject decisions, same final pose)."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from rolo_b200 import capi, synth

dp = C.POINTER(C.c_double)


def _d(a): return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def L():
    lib = capi.load()
    lib.rgt_lsq_create.restype = C.c_void_p
    lib.rgt_lsq_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
    for n in ("rgt_lsq_destroy", "rgt_lsq_begin_align", "rgt_lsq_begin_translation", "rgt_lsq_feed", "rgt_lsq_state"):
        getattr(lib, n).restype = None
    lib.rgt_lsq_begin_align.argtypes = [C.c_void_p, dp]
    lib.rgt_lsq_begin_translation.argtypes = [C.c_void_p, dp]
    lib.rgt_lsq_want.argtypes = [C.c_void_p, dp, dp]
    lib.rgt_lsq_feed.argtypes = [C.c_void_p, dp, dp, C.c_double]
    lib.rgt_lsq_state.argtypes = [C.c_void_p, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), dp]
    lib.rgt_lsq_destroy.argtypes = [C.c_void_p]
    return lib


def _run(L, h, o, guess3=None, last3=None):
    kinds = []
    pose = np.empty(16); t3 = np.empty(3)
    while True:
        k = L.rgt_lsq_want(h, _d(pose), _d(t3))
        if k == 0:
            break
        kinds.append(k)
        if k == 1:
            c, H, b = o.linearize(pose.reshape(4, 4)); L.rgt_lsq_feed(h, _d(np.ascontiguousarray(H)), _d(b), c)
        elif k == 2:
            c, H, b = o.so3_linearize(pose.reshape(4, 4)); L.rgt_lsq_feed(h, _d(np.ascontiguousarray(H)), _d(b), c)
        elif k == 3:
            L.rgt_lsq_feed(h, None, None, o.compute_error(pose.reshape(4, 4)))
        elif k == 4:
            c, H, b = o.t3_linearize(t3, guess3, last3, 0.1, 0.1); L.rgt_lsq_feed(h, _d(np.ascontiguousarray(H)), _d(b), c)
        else:
            L.rgt_lsq_feed(h, None, None, o.compute_t_error(t3, guess3, last3, 0.1, 0.1))
    return kinds


def _state(L, h):
    pose = np.empty(16); t3 = np.empty(3); it = C.c_int(); cv = C.c_int(); fl = C.c_int(); Hf = np.empty(36)
    L.rgt_lsq_state(h, _d(pose), _d(t3), C.byref(it), C.byref(cv), C.byref(fl), _d(Hf))
    return pose.reshape(4, 4), t3, it.value, bool(cv.value), bool(fl.value), Hf.reshape(6, 6)


@pytest.mark.parametrize("opt,iters,eps", [(ob.OPT_GAUSS_NEWTON, 6, 1e-30), (ob.OPT_LM, 64, None), (ob.OPT_SO3_LM, 64, None)])
def test_driver_retraces_the_oracle(L, tiny_pair, opt, iters, eps):
    tgt, src, _ = tiny_pair

    def make():
        o = ob.Oracle(); o.setNumThreads(1); o.setPolarResolution(0.175, 0.175, 2.0); o.setOptimizerType(opt); o.setMaximumIterations(iters)
        if eps:
            o.setRotationEpsilon(eps); o.setTransformationEpsilon(eps)
        o.setInputTarget(tgt); o.setInputSource(src)
        return o
    ref = make(); ref.align()
    trace = ref.getTrace()
    o = make(); o.calculateCovariances(0); o.calculateCovariances(1)
    h = C.c_void_p(L.rgt_lsq_create(opt, iters, eps or 2e-3, eps or 5e-4))
    L.rgt_lsq_begin_align(h, _d(np.eye(4).ravel()))
    kinds = _run(L, h, o)
    pose, _, it, conv, failed, Hf = _state(L, h)
    # same sequence of evaluations (oracle trace kinds: 0 linearize, 1 so3_linearize, 2 compute_error)
    assert [{1: 0, 2: 1, 3: 2}[k] for k in kinds] == list(trace[:, 0].astype(int))
    assert np.allclose(pose, ref.getFinalPoseDouble(), rtol=0, atol=1e-13)
    assert it == ref.nrIterations() and conv == ref.hasConverged() and not failed
    if opt != ob.OPT_SO3_LM:
        assert np.allclose(Hf, ref.getFinalHessian(), rtol=1e-12)
    # translation stage
    guess = pose[:3, 3].copy(); last = np.array([0.01, 0.0, 0.0])
    t_ref = ref.computeTranslation(np.zeros(3), guess, last, 0.1, 0.1, 0.3)
    o.setCtLambda(0.3)
    L.rgt_lsq_begin_translation(h, _d(np.zeros(3)))
    kinds_t = _run(L, h, o, guess, last)
    _, t3, _, _, failed, _ = _state(L, h)
    # with eps = 1e-30 the translation LM exhausts its 10 trials: both sides must break with "lm not converged!!" alike
    assert set(kinds_t) <= {4, 5} and np.allclose(t3, t_ref, rtol=0, atol=1e-13) and failed == (ref.rc == 1)
    L.rgt_lsq_destroy(h)


def test_host_math_matches_oracle(L):
    L.rgt_se3_exp.argtypes = [dp, dp]; L.rgt_ldlt_solve.argtypes = [dp, dp, C.c_int, dp]
    rng = np.random.default_rng(2)
    for scale in (1e-8, 1e-2, 1.0):
        a = rng.normal(size=6) * scale; T = np.empty(16)
        L.rgt_se3_exp(_d(a), _d(T))
        assert np.allclose(T.reshape(4, 4), ob.se3_exp(a), rtol=0, atol=1e-15)
    for n in (3, 6):
        B = rng.normal(size=(n, n)); A = np.ascontiguousarray(B @ B.T + 0.05 * np.eye(n)); b = rng.normal(size=n); x = np.empty(n)
        L.rgt_ldlt_solve(_d(A), _d(b), n, _d(x))
        assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-9)


def test_persistent_cta_schedule_covers_every_tile_once(L):
    L.rgt_build_schedule.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
    rng = np.random.default_rng(4)
    for n_pairs, grid in ((1, 296), (1, 7), (5, 296), (128, 296), (37, 64), (3, 1), (300, 296)):
        npts = rng.integers(1, 300000, size=n_pairs).astype(np.int32)
        if n_pairs == 128:
            npts[:] = 262144
        sched = np.zeros(3 * n_pairs + 1, dtype=np.int32)
        T = L.rgt_build_schedule(npts.ctypes.data_as(C.POINTER(C.c_int)), n_pairs, grid, sched.ctypes.data_as(C.POINTER(C.c_int)))
        prefix, first, count = sched[:n_pairs + 1], sched[n_pairs + 1:2 * n_pairs + 1], sched[2 * n_pairs + 1:]
        tiles = (npts.astype(np.int64) + 255) // 256
        assert T == tiles.sum() and np.array_equal(np.diff(prefix), tiles)
        G = min(grid, T) if grid > T else grid
        bounds = [(g * T) // grid for g in range(grid + 1)]
        for p in range(n_pairs):
            touching = [g for g in range(grid) if bounds[g] < prefix[p + 1] and bounds[g + 1] > prefix[p]]
            assert touching and touching[0] == first[p] and len(touching) == count[p] and touching == list(range(first[p], first[p] + count[p]))
