"""CPU tests of the product's __host__ __device__ math (lifcal_b200/csrc/lfba_math.cuh — the functions the CUDA
kernels call) compiled for the host by tests/cpu_harness: analytic Jacobian and per-track chain rule against the
oracle's autodiff and against the golden vectors made by the reference's own headers. Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers
from lifcal_b200 import capi
from oracle import binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "functor_kat.npz")


@pytest.fixture(scope="module")
def harness(built):
    src = os.path.join(HERE, "cpu_harness", "harness.cpp")
    out = os.path.join(HERE, "cpu_harness", "libharness.so")
    dep = os.path.join(HERE, "..", "lifcal_b200", "csrc", "lfba_math.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-std=c++17", "-shared", "-x", "c++", "-o", out, src])
    H = C.CDLL(out)
    dp = capi.c_double_p
    H.harness_eval.argtypes = [C.POINTER(capi.Problem), dp, dp, dp, dp, dp, dp, dp]
    H.harness_spd3_inverse.argtypes = [dp, dp]
    H.harness_distance.argtypes = [dp, dp, C.c_double, C.c_double, dp, dp]
    return H


def _eval(H, pa, cam, vw, pt):
    n = pa.n_obs
    res, jc, jv, jp = np.zeros(2 * n), np.zeros((n, 2, 17)), np.zeros((n, 2, 6)), np.zeros((n, 2, 3))
    p = pa.as_struct()
    assert H.harness_eval(C.byref(p), capi._dp(np.ascontiguousarray(cam)), capi._dp(np.ascontiguousarray(vw)),
                          capi._dp(np.ascontiguousarray(pt)), capi._dp(res), capi._dp(jc), capi._dp(jv), capi._dp(jp)) == 0
    return res.reshape(n, 2), jc, jv, jp


def _colrel(a, b):
    worst = 0.0
    for c in range(a.shape[-1]):
        den = np.max(np.abs(b[..., c]))
        if den > 0:
            worst = max(worst, np.max(np.abs(a[..., c] - b[..., c])) / den)
        else:
            assert np.max(np.abs(a[..., c])) == 0
    return worst


def test_analytic_jacobian_matches_reference_golden(harness):
    g = np.load(GOLD)
    spx, scale = float(g["spx"][0]), float(g["scale"][0])
    for mc in helpers.all_model_configs():
        for ab in (capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS, capi.CFG_REFINE_POSES):
            cfg = mc | ab
            for ci in range(2):
                key = f"cfg{cfg:#06x}_{ci}"
                obs, ml = g[key + "_obs"], g[key + "_ml"]
                n = obs.shape[0]
                pa = capi.ProblemArrays(cfg, 0, spx, spx, scale, n, n, obs[:, 0], obs[:, 1], ml[:, 0], ml[:, 1],
                                        np.arange(n), np.arange(n))
                res, jc, jv, jp = _eval(harness, pa, g[key + "_camera"], g[key + "_views"].ravel(), g[key + "_points"].ravel())
                assert np.max(np.abs(res - g[key + "_res"])) < 1e-10, key
                assert _colrel(jc, g[key + "_jc"]) < 1e-12, key
                assert _colrel(jv, g[key + "_jv"]) < 1e-12, key
                if ab & capi.CFG_REFINE_POINTS:
                    assert _colrel(jp, g[key + "_jp"]) < 1e-12, key


def test_analytic_jacobian_matches_oracle_autodiff_random(harness):
    rng = np.random.default_rng(42)
    for mc in helpers.all_model_configs():
        cfg = mc | capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS
        b = helpers.random_blocks(rng, cfg, 300, signs=True)
        n = 300
        pa = capi.ProblemArrays(cfg, 0, b["spx"], b["spy"], b["scale"], n, n, b["obs"][:, 0], b["obs"][:, 1],
                                b["ml"][:, 0], b["ml"][:, 1], np.arange(n), np.arange(n))
        o = ob.evaluate(pa, b["cams"][0], b["views"].ravel(), b["points"].ravel())
        res, jc, jv, jp = _eval(harness, pa, b["cams"][0], b["views"].ravel(), b["points"].ravel())
        assert np.max(np.abs(res - o["residuals"])) < 1e-10
        assert _colrel(jc, o["jac_camera"]) < 1e-12
        assert _colrel(jv, o["jac_view"]) < 1e-12
        assert _colrel(jp, o["jac_point"]) < 1e-12


def test_feature_gram_forms_reproduce_jacobian_products(harness):
    """The fused kernel never forms the Jacobian: it accumulates the weighted Gram matrix of NC two-component features
    and expands it per track (obs_features9 + gram9_expand + GramMap). On single observations that expansion must
    equal w * (Jacobian block products) up to rounding — for every model-flag combination, random poses/points."""
    harness.harness_feature_worst.restype = C.c_double
    harness.harness_feature9_worst.restype = C.c_double
    harness.harness_feature_reset()
    rng = np.random.default_rng(7)
    for mc in helpers.all_model_configs():
        cfg = mc | capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS
        b = helpers.random_blocks(rng, cfg, 200, signs=True)
        n = 200
        pa = capi.ProblemArrays(cfg, 0, b["spx"], b["spy"], b["scale"], n, n, b["obs"][:, 0], b["obs"][:, 1],
                                b["ml"][:, 0], b["ml"][:, 1], np.arange(n), np.arange(n))
        _eval(harness, pa, b["cams"][0], b["views"].ravel(), b["points"].ravel())
    assert harness.harness_feature9_worst() < 1e-9, harness.harness_feature9_worst()


def test_spd3_inverse_and_distance(harness):
    rng = np.random.default_rng(0)
    for _ in range(100):
        M = rng.standard_normal((3, 3))
        A = M @ M.T + 1e-3 * np.eye(3)
        a6 = np.array([A[0, 0], A[0, 1], A[0, 2], A[1, 1], A[1, 2], A[2, 2]])
        inv6 = np.zeros(6)
        assert harness.harness_spd3_inverse(capi._dp(a6), capi._dp(inv6)) == 1
        Ai = np.array([[inv6[0], inv6[1], inv6[2]], [inv6[1], inv6[3], inv6[4]], [inv6[2], inv6[4], inv6[5]]])
        assert np.allclose(Ai @ A, np.eye(3), atol=1e-9)
    bad = np.array([1.0, 2.0, 0.0, 1.0, 0.0, 1.0])  # indefinite
    assert harness.harness_spd3_inverse(capi._dp(bad), capi._dp(np.zeros(6))) == 0
    g = np.load(GOLD)
    for i in range(g["dc_p1"].shape[0]):
        r = C.c_double()
        j = np.zeros(3)
        harness.harness_distance(capi._dp(np.ascontiguousarray(g["dc_p1"][i])), capi._dp(np.ascontiguousarray(g["dc_p2"][i])),
                                 float(g["dc_dist"][i]), float(g["dc_sigma"][i]), C.byref(r), capi._dp(j))
        assert abs(r.value - g["dc_res"][i]) <= 1e-12 * max(1.0, abs(g["dc_res"][i]))
        assert np.allclose(j, g["dc_jac"][i, :3], rtol=1e-12, atol=1e-14)
        assert np.allclose(-j, g["dc_jac"][i, 3:], rtol=1e-12, atol=1e-14)
