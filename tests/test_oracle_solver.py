"""CPU tests (no GPU): the oracle's Ceres-2.1.0-equivalent LM loop.
 - reproduces the committed iteration tables (tests/golden/solver_small.json, generated with the REFERENCE
   functor plugged into the loop) with its own restated functor
 - Ceres trust-region invariants on the logged rows
The LM loop itself is "parity unpinned" (Ceres is not available here); these tests pin it against regressions
and against the documented semantics."""
import json
import os

import numpy as np
import pytest

from lifcal_b200 import capi
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(__file__), "golden", "solver_small.json")


def _cases():
    with open(GOLD) as f:
        return json.load(f)


@pytest.mark.parametrize("name", list(_cases().keys()))
def test_solver_reproduces_golden_tables(built, name):
    case = _cases()[name]
    sc = capi.make_scene(None, **case["scene"])
    assert sc.problem.n_obs == case["n_obs"]
    cam, vw, pt, s = ob.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init, threads=2)
    assert s["num_iterations"] == case["num_iterations"]
    assert s["stop_reason"] == case["stop_reason"]
    for r, gr in zip(s["iterations"], case["rows"]):
        assert r["iteration"] == gr["iteration"]
        assert r["step_is_successful"] == gr["step_is_successful"]
        assert abs(r["cost"] - gr["cost"]) <= 1e-9 * abs(gr["cost"])
        assert abs(r["trust_region_radius"] - gr["trust_region_radius"]) <= 1e-6 * gr["trust_region_radius"]
    assert abs(s["final_cost"] - case["final_cost"]) <= 1e-9 * case["final_cost"]
    assert np.allclose(cam, case["camera"], rtol=1e-7, atol=1e-12)
    assert np.allclose(vw, case["views"], rtol=1e-6, atol=1e-9)
    assert np.allclose(pt[:30], case["points_head"], rtol=1e-7, atol=1e-7)


def test_trust_region_invariants(built):
    sc = capi.make_scene(None, n_points=80, n_frames=5, n_constraints=2, seed=99, init_intrinsics_rel=2e-3)
    cam, vw, pt, s = ob.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init, threads=2)
    rows = s["iterations"]
    assert rows[0]["iteration"] == 0 and rows[0]["trust_region_radius"] == 1e4
    assert s["initial_cost"] == rows[0]["cost"]
    cost = rows[0]["cost"]
    radius = 1e4
    dec = 2.0
    for r in rows[1:]:
        if r["step_is_successful"]:
            assert r["cost"] < cost
            assert r["relative_decrease"] > 1e-3
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * r["relative_decrease"] - 1.0) ** 3))
            dec = 2.0
            cost = r["cost"]
        else:
            radius /= dec
            dec *= 2.0
        assert abs(r["trust_region_radius"] - radius) <= 1e-12 * radius
    assert abs(s["final_cost"] - cost) <= 1e-15 * cost
    # unmodified inputs, outputs differ
    assert not np.array_equal(cam, sc.camera_init)


def test_invalid_flag_combination_is_rejected(built):
    # refinePoses = 0 with refine3Dpoints = 1 null-derefs in the reference
    sc = capi.make_scene(None, n_points=10, n_frames=2, seed=1)
    pa = sc.problem.with_config(2 | capi.CFG_REFINE_POINTS)
    cam, vw, pt, s = ob.solve(pa, sc.camera_init, sc.views_init, sc.points_init)
    assert s["status"] == capi.INVALID_ARGUMENT


def test_recalib_holds_f_and_B_and_respects_bounds(built):
    sc = capi.make_scene(None, n_points=60, n_frames=5, seed=13, calib_type=capi.RECALIBRATION)
    cam, vw, pt, s = ob.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init, threads=2)
    assert cam[0] == sc.camera_init[0] and cam[2] == sc.camera_init[2]  # SubsetManifold(17,{0,2})
    for j in (1, 3, 4):
        assert 0.7 * sc.camera_init[j] <= cam[j] <= 1.3 * sc.camera_init[j]
    assert s["final_cost"] < s["initial_cost"]


def test_scene_generator_is_deterministic_and_shardable(built):
    a = capi.make_scene(None, n_points=200, n_frames=8, window=3, seed=5, order=1)
    b = capi.make_scene(None, n_points=200, n_frames=8, window=3, seed=5, order=1, num_threads=1)
    assert np.array_equal(a.problem.obs_x, b.problem.obs_x) and np.array_equal(a.problem.point_idx, b.problem.point_idx)
    lo = capi.make_scene(None, n_points=200, n_frames=8, window=3, seed=5, order=1, point_begin=0, point_end=120)
    hi = capi.make_scene(None, n_points=200, n_frames=8, window=3, seed=5, order=1, point_begin=120, point_end=200)
    assert lo.problem.n_obs + hi.problem.n_obs == a.problem.n_obs
    assert np.array_equal(np.concatenate([lo.problem.obs_y, hi.problem.obs_y]), a.problem.obs_y)
    # frame-major order holds the same multiset of observations
    c = capi.make_scene(None, n_points=200, n_frames=8, window=3, seed=5, order=0)
    assert np.all(np.diff(c.problem.frame_idx) >= 0)
    ka = np.lexsort((a.problem.ml_y, a.problem.ml_x, a.problem.frame_idx, a.problem.point_idx))
    kc = np.lexsort((c.problem.ml_y, c.problem.ml_x, c.problem.frame_idx, c.problem.point_idx))
    assert np.array_equal(a.problem.obs_x[ka], c.problem.obs_x[kc])
    # float32 round trip of observations and lens centres (src/CameraCalibration.cpp:748-762)
    assert np.array_equal(a.problem.obs_x, a.problem.obs_x.astype(np.float32).astype(np.float64))
    assert np.array_equal(a.problem.ml_x, a.problem.ml_x.astype(np.float32).astype(np.float64))


def test_oracle_minimum_agrees_with_an_independent_solver(built):
    """Ceres is not available here, so the LM loop of the oracle is 'parity unpinned'. An independent check that it at
    least lands in the right place: scipy's trust-region least squares on the SAME residuals/Jacobians (non-robust
    cost, so that both minimise 0.5 sum r^2) must reach the same minimum cost. This does not pin Ceres' iteration
    semantics (those are covered by the invariants above), only the fixed point."""
    from scipy.optimize import least_squares
    cfg = 2 | capi.CFG_TANGENTIAL | capi.CFG_MLADJ | capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS  # no robust loss
    sc = capi.make_scene(None, n_points=40, n_frames=4, seed=123, config=cfg)
    pa = sc.problem
    F, P, n = pa.n_frames, pa.n_points, pa.n_obs
    live = 9

    def unpack(x):
        cam = sc.camera_init.copy()
        cam[:live] = x[:live]
        return cam, x[live:live + 6 * F].copy(), x[live + 6 * F:].copy()

    def fun(x):
        cam, vw, pt = unpack(x)
        return ob.evaluate(pa, cam, vw, pt, jacobians=False)["residuals"].ravel()

    def jac(x):
        cam, vw, pt = unpack(x)
        e = ob.evaluate(pa, cam, vw, pt)
        J = np.zeros((2 * n, live + 6 * F + 3 * P))
        rows = np.arange(2 * n).reshape(n, 2)
        J[:, :live] = e["jac_camera"].reshape(2 * n, 17)[:, :live]
        for i in range(n):
            f, p = pa.frame_idx[i], pa.point_idx[i]
            J[rows[i], live + 6 * f:live + 6 * f + 6] = e["jac_view"][i]
            J[rows[i], live + 6 * F + 3 * p:live + 6 * F + 3 * p + 3] = e["jac_point"][i]
        return J

    x0 = np.concatenate([sc.camera_init[:live], sc.views_init, sc.points_init])
    ref = least_squares(fun, x0, jac=jac, method="trf", x_scale="jac", ftol=1e-14, xtol=1e-14, gtol=1e-14, max_nfev=200)
    o = ob.default_options(function_tolerance=1e-13, parameter_tolerance=1e-13, max_num_iterations=200)
    _, _, _, s = ob.solve(pa, sc.camera_init, sc.views_init, sc.points_init, options=o, threads=2)
    assert ref.cost > 0
    assert abs(s["final_cost"] - ref.cost) <= 1e-6 * ref.cost, (s["final_cost"], ref.cost)
    # with the reference's own tolerances (ftol 1e-6) the oracle stops slightly above that minimum, never below it
    _, _, _, s2 = ob.solve(pa, sc.camera_init, sc.views_init, sc.points_init, threads=2)
    assert ref.cost * (1 - 1e-9) <= s2["final_cost"] <= ref.cost * (1 + 1e-3)


def test_oracle_robust_minimum_agrees_with_an_independent_solver(built):
    """The same independent check for the ROBUST cost 0.5 sum rho(|r_i|^2), rho = CauchyLoss(0.5) per 2-vector block
    (src/CameraCalibration.cpp:905): scipy minimises 0.5 sum |g(s_i) r_i|^2 with g = sqrt(rho(s) / s), which is the very
    same function, on residuals and Jacobians of the REFERENCE's own functor (oracle/_ref). Ceres' Corrector (incl. its
    second-order term) only shapes the path; the minimum it must reach is this one."""
    from scipy.optimize import least_squares
    cfg = 2 | capi.CFG_TANGENTIAL | capi.CFG_MLADJ | capi.CFG_ROBUST | capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS
    sc = capi.make_scene(None, n_points=40, n_frames=4, seed=321, config=cfg)
    pa = sc.problem
    F, P, n = pa.n_frames, pa.n_points, pa.n_obs
    live, b = 9, 0.25  # a = 0.5: rho(s) = b log(1 + s / b)

    def unpack(x):
        cam = sc.camera_init.copy()
        cam[:live] = x[:live]
        return cam, x[live:live + 6 * F].copy(), x[live + 6 * F:].copy()

    def g_and_dg(sq):
        sq = np.maximum(sq, 1e-300)
        h = b * np.log1p(sq / b) / sq
        dh = (1.0 / (1.0 + sq / b) * sq - b * np.log1p(sq / b)) / (sq * sq)
        small = sq < 1e-8  # series: h = 1 - s / (2 b) + ...
        h = np.where(small, 1.0 - sq / (2 * b), h)
        dh = np.where(small, -1.0 / (2 * b), dh)
        g = np.sqrt(h)
        return g, dh / (2.0 * g)

    def fun(x):
        cam, vw, pt = unpack(x)
        r = ob.evaluate(pa, cam, vw, pt, jacobians=False, use_ref=True)["residuals"]
        g, _ = g_and_dg(np.sum(r * r, axis=1))
        return (g[:, None] * r).ravel()

    def jac(x):
        cam, vw, pt = unpack(x)
        e = ob.evaluate(pa, cam, vw, pt, use_ref=True)
        r = e["residuals"]
        g, dg = g_and_dg(np.sum(r * r, axis=1))
        J = np.zeros((2 * n, live + 6 * F + 3 * P))
        rows = np.arange(2 * n).reshape(n, 2)
        for i in range(n):
            f, p = pa.frame_idx[i], pa.point_idx[i]
            Ji = np.zeros((2, J.shape[1]))
            Ji[:, :live] = e["jac_camera"][i][:, :live]
            Ji[:, live + 6 * f:live + 6 * f + 6] = e["jac_view"][i]
            Ji[:, live + 6 * F + 3 * p:live + 6 * F + 3 * p + 3] = e["jac_point"][i]
            J[rows[i]] = (g[i] * np.eye(2) + 2.0 * dg[i] * np.outer(r[i], r[i])) @ Ji  # d (g(s) r) = (g I + 2 g' r r^T) dr
        return J

    x0 = np.concatenate([sc.camera_init[:live], sc.views_init, sc.points_init])
    ref = least_squares(fun, x0, jac=jac, method="trf", x_scale="jac", ftol=1e-14, xtol=1e-14, gtol=1e-14, max_nfev=300)
    # the transformed problem IS the robust cost: check against the oracle's own cost function at the start point
    e0 = ob.evaluate(pa, sc.camera_init, sc.views_init, sc.points_init, jacobians=False)
    assert abs(0.5 * np.sum(fun(x0) ** 2) - e0["cost"]) <= 1e-12 * e0["cost"]
    o = ob.default_options(function_tolerance=1e-13, parameter_tolerance=1e-13, max_num_iterations=300)
    _, _, _, s = ob.solve(pa, sc.camera_init, sc.views_init, sc.points_init, options=o, threads=2)
    assert ref.cost > 0 and ref.cost < 0.9 * e0["cost"]
    assert abs(s["final_cost"] - ref.cost) <= 1e-6 * ref.cost, (s["final_cost"], ref.cost)


@pytest.mark.parametrize("scene", [
    dict(n_points=200, n_frames=6, n_constraints=2, seed=7),                       # constraints + coupled points
    dict(n_points=300, n_frames=24, window=4, seed=5, order=1),                    # windowed (cfg4 family)
    dict(n_points=120, n_frames=5, seed=13, calib_type=capi.RECALIBRATION),        # manifold + bounds + line search
])
def test_streaming_mode_equals_stored_mode(built, scene):
    """The Jacobian-free (block-recompute) mode that makes the 1M x 1000 scene fit the host follows the same
    algorithm as the stored-Jacobian mode: same rows, same decisions, costs equal to rounding."""
    sc = capi.make_scene(None, **scene)
    init = (sc.camera_init, sc.views_init, sc.points_init)
    a = ob.solve(sc.problem, *init, threads=2)
    b = ob.solve(sc.problem, *init, threads=2, streaming=True)
    sa, sb = a[3], b[3]
    assert sa["num_iterations"] == sb["num_iterations"] and sa["stop_reason"] == sb["stop_reason"]
    assert sb["block_passes"] >= 3 * (sb["num_iterations"] - 1) + 1  # evaluate + eliminate + back-substitute
    for ra, rb in zip(sa["iterations"], sb["iterations"]):
        assert ra["step_is_successful"] == rb["step_is_successful"]
        assert abs(ra["cost"] - rb["cost"]) <= 1e-12 * abs(ra["cost"])
        assert abs(ra["trust_region_radius"] - rb["trust_region_radius"]) <= 1e-6 * ra["trust_region_radius"]
        assert abs(ra["relative_decrease"] - rb["relative_decrease"]) <= 1e-6
    assert abs(sa["final_cost"] - sb["final_cost"]) <= 1e-12 * sa["final_cost"]
    assert np.allclose(a[0], b[0], rtol=1e-7, atol=1e-12)
