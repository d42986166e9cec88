"""GPU tests (-m gpu) of lfba_project_to_raw against the CPU oracle: bit-exact lens selection and
order, float32-exact coordinates; and consistency of its output with the projection model of the solve."""
import numpy as np
import pytest

import helpers
from lifcal_b200 import api, capi
from oracle import binding as ob

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built):
    if api.device_count() < 1:
        pytest.fail("no CUDA device: GPU tests cannot fall back to anything")
    return True


@pytest.mark.parametrize("rotation,on", [(0.003, True), (0.0, False), (-0.02, True)])
def test_device_projection_is_bit_identical_to_the_oracle(gpu, rotation, on):
    g = helpers.make_lens_grid(raw=640, rotation=rotation)
    g.rotation_on_grid = int(on)
    rng = np.random.default_rng(3)
    m = 20000
    fx = rng.random(m) * (640 / 2 - 1)
    fy = rng.random(m) * (640 / 2 - 1)
    vd = 1.0 + rng.random(m) * 21.5          # includes depths outside (2, 20): dropped (:655)
    vd[::97] = 2.0                            # boundary values are excluded too
    vd[5::101] = 20.0
    fx[7::113] = 0.0                          # image corners / borders: clamped lens lookups (:729-732)
    fy[9::127] = 640 / 2 - 1
    frames = np.sort(rng.integers(0, 12, m)).astype(np.int32)   # frame-major feature list
    points = rng.integers(0, 5000, m).astype(np.int32)
    o = ob.project_to_raw(g, fx, fy, vd)
    d = api.project_to_raw(g, fx, fy, vd, frames, points)
    assert d["obs_x"].size == o["obs_x"].size > 100000
    for k in ("obs_x", "obs_y", "ml_x", "ml_y"):
        assert np.array_equal(d[k], o[k]), k          # same lenses, same order, float32-exact coordinates
    assert np.array_equal(d["point_idx"], points[o["feature"]])
    assert np.array_equal(d["frame_idx"], frames[o["feature"]])
    assert np.all(np.diff(d["frame_idx"]) >= 0)      # the concatenation of the reference's per-frame lists


def test_empty_and_invalid_inputs(gpu):
    g = helpers.make_lens_grid(raw=256)
    d = api.project_to_raw(g, np.zeros(0), np.zeros(0), np.zeros(0))
    assert d["obs_x"].size == 0
    d = api.project_to_raw(g, [10.0, 20.0], [10.0, 20.0], [1.0, 25.0])  # no valid virtual depth
    assert d["obs_x"].size == 0
    g.map_next[:] = -1                                                   # mapNextMl == NULL everywhere ("Error." :672)
    d = api.project_to_raw(g, [10.0], [10.0], [5.0])
    assert d["obs_x"].size == 0


def test_generated_observations_fit_the_projection_model_of_the_solve(gpu):
    """Features rendered from a synthetic scene's ground truth (thin-lens virtual image + virtual depth)
    -> lfba_project_to_raw -> lfba_eval at the ground truth: without distortion the plenoptic projection of the solve
    reproduces every generated micro-image point to float32 rounding, so observation generation and the residual functor
    agree on geometry, lens centres and indices."""
    cfg = capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS
    sc = capi.make_scene(None, n_points=400, n_frames=4, seed=3, config=cfg)
    cam, views, pts = sc.camera_true, sc.views_true.reshape(-1, 6), sc.points_true.reshape(-1, 3)
    fL, bL0, B, cx, cy = cam[:5]
    raw, scale = 2048, 2
    g = helpers.make_lens_grid(raw=raw, rotation=0.003, scale=scale)
    s_tot = sc.problem.spx  # total-focus pixel size (mm)
    fxs, fys, vds, fis, pis = [], [], [], [], []
    for f in range(views.shape[0]):
        R = helpers_rot(views[f, :3])
        pc = pts @ R.T + views[f, 3:]
        Z = pc[:, 2]
        bL = fL * Z / (Z - fL)
        v = (bL - bL0) / B
        xv = pc[:, 0] / Z * bL / s_tot + cx
        yv = pc[:, 1] / Z * bL / s_tot + cy
        ok = (v > 2.2) & (v < 19) & (xv > 40) & (xv < raw / scale - 40) & (yv > 40) & (yv < raw / scale - 40)
        fxs.append(xv[ok]); fys.append(yv[ok]); vds.append(v[ok])
        fis.append(np.full(ok.sum(), f, np.int32)); pis.append(np.where(ok)[0].astype(np.int32))
    fx, fy, vd = map(np.concatenate, (fxs, fys, vds))
    fi, pi = np.concatenate(fis), np.concatenate(pis)
    assert fx.size > 200
    d = api.project_to_raw(g, fx, fy, vd, fi, pi)
    assert d["obs_x"].size > 10 * fx.size
    pa = capi.ProblemArrays(cfg, 0, sc.problem.spx, sc.problem.spy, float(scale), views.shape[0], pts.shape[0], d["obs_x"],
                            d["obs_y"], d["ml_x"], d["ml_y"], d["point_idx"], d["frame_idx"])
    ev = api.evaluate(pa, cam, sc.views_true, sc.points_true, jacobians=False)
    # float32 features (x, y, v rounded to float) and float32 arithmetic: |r| ~ 1e-4 px; a wrong lens / index / sign is > 1 px
    assert np.max(np.abs(ev["residuals"])) < 5e-3, np.max(np.abs(ev["residuals"]))
    assert ev["stats"]["num_inliers"] == d["obs_x"].size


def helpers_rot(a):
    cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def test_linear_initialisation_on_the_device_matches_the_svd_oracle(gpu):
    # SURVEY.md 8(f) N4, src/CameraCalibration.cpp:456-498: bL = v B + bL0 over all (frame, feature) pairs
    from lifcal_b200 import init_params as ip
    sc = capi.make_scene(None, n_points=3000, n_frames=12, seed=9)
    cam, views, pts = sc.camera_true, sc.views_true.reshape(-1, 6), sc.points_true.reshape(-1, 3)
    fL, bL0, B = cam[0], cam[1], cam[2]
    rng = np.random.default_rng(2)
    fi = np.repeat(np.arange(12, dtype=np.int32), 3000)
    pi = np.tile(np.arange(3000, dtype=np.int32), 12)
    z = np.zeros(fi.size)
    for f in range(12):
        z[fi == f] = (pts @ helpers_rot(views[f, :3]).T + views[f, 3:])[:, 2]
    v = (fL * z / (z - fL) - bL0) / B + 2e-3 * rng.standard_normal(z.size)
    v[:40] = 1.7                                   # rejected rows: v < 2
    pts2 = pts.copy()
    pts2[5] = -views[0, 3:] @ helpers_rot(views[0, :3]) + np.array([0, 0, 10.0]) @ helpers_rot(views[0, :3])  # Z = 10 < fL: bL < 0
    sp = sc.problem.spx
    got = ip.init_plenoptic_parameters(fL / sp, sp, v, fi, pi, views, pts2)
    ref = ob.init_plenoptic(fL / sp, sp, v, fi, pi, views, pts2)
    assert got[0] == ref[0]
    assert abs(got[1] - ref[1]) <= 1e-11 * abs(ref[1]) and abs(got[2] - ref[2]) <= 1e-11 * abs(ref[2])
    assert abs(got[1] - B) < 2e-2 and abs(got[2] - bL0) < 2e-1  # recovers the camera it was rendered from
