"""oracle/binding.py — TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/liblfba_oracle.so) and of
the reference functors compiled in place (oracle/_ref/libref_functor.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (lifcal_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from lifcal_b200 import capi  # noqa: E402  (struct mirrors only)

c_double_p = C.POINTER(C.c_double)
BLOCK_FN = C.CFUNCTYPE(C.c_int, C.c_uint32, c_double_p, c_double_p, C.c_double, C.c_double, C.c_double,
                       c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p)

_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liblfba_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built (make -C oracle)")
        L = C.CDLL(path)
        L.oracle_eval.argtypes = [C.POINTER(capi.Problem), c_double_p, c_double_p, c_double_p, c_double_p,
                                  c_double_p, c_double_p, c_double_p, c_double_p, C.POINTER(capi.ReprojStats),
                                  C.c_double, C.c_int, C.c_void_p]
        L.oracle_solve.argtypes = [C.POINTER(capi.Problem), C.POINTER(capi.Options), c_double_p, c_double_p,
                                   c_double_p, C.POINTER(capi.Summary), C.c_int, C.c_double, C.c_void_p]
        L.oracle_solve_streaming.argtypes = L.oracle_solve.argtypes
        L.oracle_time_eval.argtypes = [C.POINTER(capi.Problem), c_double_p, c_double_p, c_double_p, C.c_int,
                                       C.c_int, c_double_p]
        L.oracle_max_threads.restype = C.c_int
        i32p, f32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_int64)
        L.oracle_epi_web.argtypes = [C.c_float, C.c_float, C.c_int, i32p, i32p, c_double_p, i32p]
        L.oracle_epi_make.argtypes = [C.c_double, C.c_double, C.c_double, c_double_p]
        L.oracle_epi_add.argtypes = [c_double_p, c_double_p, c_double_p]
        L.oracle_project_to_raw.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int32,
                                            f32p, f32p, i32p, i32p, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_int64,
                                            c_double_p, c_double_p, c_double_p, c_double_p, i64p]
        L.oracle_project_to_raw.restype = C.c_int64
        _lib = L
    return _lib


def ref_lib():
    """The reference's own functors (None if oracle/_ref was not built)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libref_functor.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_block_eval.argtypes = [C.c_uint32, c_double_p, c_double_p, C.c_double, C.c_double, C.c_double,
                                     c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                     c_double_p]
        R.ref_block_eval.restype = C.c_int
        R.ref_distance_eval.argtypes = [C.c_double, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p]
        R.ref_distance_eval.restype = C.c_int
        R.ref_project_point.argtypes = [c_double_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                        c_double_p, c_double_p, c_double_p, C.c_int, c_double_p, C.c_int,
                                        c_double_p]
        R.ref_pose_matrix.argtypes = [c_double_p, c_double_p]
        if hasattr(R, "ref_epi_make"):
            R.ref_epi_make.argtypes = [C.c_double, C.c_double, C.c_double, c_double_p]
            R.ref_epi_add.argtypes = [c_double_p, c_double_p, c_double_p]
        _ref = R
    return _ref


def ref_block_fn_ptr():
    R = ref_lib()
    if R is None:
        return None
    return C.cast(R.ref_block_eval, C.c_void_p)


def default_options(**kw) -> capi.Options:
    """The reference's solver options (src/CameraCalibration.cpp:955-962) over Ceres 2.1.0 defaults."""
    o = capi.Options()
    o.max_num_iterations = 200
    o.function_tolerance = 1e-6
    o.parameter_tolerance = 1e-8
    o.gradient_tolerance = 1e-10
    o.initial_trust_region_radius = 1e4
    o.max_trust_region_radius = 1e16
    o.min_trust_region_radius = 1e-32
    o.min_relative_decrease = 1e-3
    o.min_lm_diagonal = 1e-6
    o.max_lm_diagonal = 1e32
    o.max_num_consecutive_invalid_steps = 5
    o.loss_scale = 0.5
    o.minimizer_progress_to_stdout = 0
    o.device = -1
    o.num_gpus = 1
    o.profile = 0
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def evaluate(pa: capi.ProblemArrays, camera, views, points, jacobians=True, threads=0, use_ref=False,
             inlier_threshold=1.0):
    L = lib()
    n = pa.n_obs
    cam = np.ascontiguousarray(camera, np.float64)
    vw = np.ascontiguousarray(views, np.float64)
    pt = np.ascontiguousarray(points, np.float64)
    res = np.zeros(2 * n)
    jc = np.zeros((n, 2, 17)) if jacobians else None
    jv = np.zeros((n, 2, 6)) if jacobians else None
    jp = np.zeros((n, 2, 3)) if jacobians else None
    cost = C.c_double(0)
    st = capi.ReprojStats()
    p = pa.as_struct()
    fn = ref_block_fn_ptr() if use_ref else None
    rc = L.oracle_eval(C.byref(p), capi._dp(cam), capi._dp(vw), capi._dp(pt), capi._dp(res), capi._dp(jc),
                       capi._dp(jv), capi._dp(jp), C.byref(cost), C.byref(st), inlier_threshold, threads, fn)
    if rc != 0:
        raise RuntimeError(f"oracle_eval rc={rc}")
    return {"residuals": res.reshape(n, 2), "jac_camera": jc, "jac_view": jv, "jac_point": jp,
            "cost": cost.value,
            "stats": {"std_x": st.std_x, "std_y": st.std_y, "mae_x": st.mae_x, "mae_y": st.mae_y,
                      "num_points": st.num_points, "num_inliers": st.num_inliers}}


def solve(pa: capi.ProblemArrays, camera, views, points, options=None, threads=0, max_seconds=0.0,
          use_ref=False, streaming=False):
    """Returns (camera, views, points, summary_dict); inputs are not modified. streaming=True: Jacobian-free
    block-recompute mode (oracle_solve_streaming) for scenes whose stored Jacobian would not fit the host."""
    L = lib()
    cam = np.array(camera, np.float64, copy=True)
    vw = np.array(views, np.float64, copy=True)
    pt = np.array(points, np.float64, copy=True)
    o = options if options is not None else default_options()
    s, rows = capi.new_summary(max(8, o.max_num_iterations + 8))
    p = pa.as_struct()
    fn = ref_block_fn_ptr() if use_ref else None
    entry = L.oracle_solve_streaming if streaming else L.oracle_solve
    rc = entry(C.byref(p), C.byref(o), capi._dp(cam), capi._dp(vw), capi._dp(pt), C.byref(s), threads,
               max_seconds, fn)
    d = capi.summary_to_dict(s, rows)
    d["status"] = rc
    d["block_passes"] = d.pop("num_lenses") if streaming else 0
    return cam, vw, pt, d


def time_eval(pa: capi.ProblemArrays, camera, views, points, reps=1, threads=0) -> float:
    L = lib()
    sec = C.c_double(0)
    p = pa.as_struct()
    cam = np.ascontiguousarray(camera, np.float64)
    vw = np.ascontiguousarray(views, np.float64)
    pt = np.ascontiguousarray(points, np.float64)
    rc = L.oracle_time_eval(C.byref(p), capi._dp(cam), capi._dp(vw), capi._dp(pt), reps, threads, C.byref(sec))
    if rc != 0:
        raise RuntimeError(f"oracle_time_eval rc={rc}")
    return sec.value


def max_threads() -> int:
    return int(lib().oracle_max_threads())


# ---- N2: epipolar web + projection of total-focus features into the micro images (oracle/project_raw.hpp) ----
def epi_web(lens_diameter, rotation=0.0, rotation_on_grid=False):
    L = lib()
    nl, ng = C.c_int32(0), C.c_int32(0)
    L.oracle_epi_web(float(lens_diameter), float(rotation), int(bool(rotation_on_grid)), C.byref(nl), C.byref(ng), None, None)
    lines = np.zeros((nl.value, 3))
    gb = np.zeros(ng.value + 1, np.int32)
    L.oracle_epi_web(float(lens_diameter), float(rotation), int(bool(rotation_on_grid)), C.byref(nl), C.byref(ng),
                     capi._dp(lines), capi._ip(gb))
    return lines, gb


def project_to_raw(grid: "capi.LensGrid", feat_x, feat_y, vdepth):
    """Returns obs_x, obs_y, ml_x, ml_y and the feature index of every observation, in the reference's order."""
    L = lib()
    fx = np.ascontiguousarray(feat_x, np.float64)
    fy = np.ascontiguousarray(feat_y, np.float64)
    vd = np.ascontiguousarray(vdepth, np.float64)
    f32p, i32p, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    args = [grid.raw_width, grid.raw_height, grid.scale, float(grid.lens_diameter), float(grid.lens_validity_radius_2),
            float(grid.rotation), grid.rotation_on_grid, int(grid.lens_cx.size), grid.lens_cx.ctypes.data_as(f32p),
            grid.lens_cy.ctypes.data_as(f32p), capi._ip(grid.map_next), capi._ip(grid.map_ml), fx.size, capi._dp(fx),
            capi._dp(fy), capi._dp(vd)]
    n = L.oracle_project_to_raw(*args, 0, None, None, None, None, None)
    out = {k: np.zeros(n) for k in ("obs_x", "obs_y", "ml_x", "ml_y")}
    feat = np.zeros(n, np.int64)
    if n:
        L.oracle_project_to_raw(*args, n, capi._dp(out["obs_x"]), capi._dp(out["obs_y"]), capi._dp(out["ml_x"]),
                                capi._dp(out["ml_y"]), feat.ctypes.data_as(i64p))
    out["feature"] = feat
    return out


# ---- N4: linear initialisation (numpy restatement of src/CameraCalibration.cpp:456-498) ----
def init_plenoptic(fph_init, pixel_size_totfoc, vdepth, frame_idx, point_idx, views, points):
    """a = [v 1] (N x 2), b = bL = fL Z / (Z - fL) with Z = (worldToCam X).z; rows with v < 2 or bL < 0 are zeroed in a AND
    b (:483-488); x = thin-SVD least squares (:492, numpy lstsq is SVD-based). Returns (fL_init, B_init, bL0_init)."""
    fL = float(fph_init) * float(pixel_size_totfoc)                     # :460
    vw = np.asarray(views, np.float64).reshape(-1, 6)
    pt = np.asarray(points, np.float64).reshape(-1, 3)
    v = np.asarray(vdepth, np.float64)
    z = np.zeros(v.size)
    for f in np.unique(frame_idx):
        a0, a1, a2 = vw[f, :3]
        cx, sx, cy, sy, cz, sz = np.cos(a0), np.sin(a0), np.cos(a1), np.sin(a1), np.cos(a2), np.sin(a2)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        sel = np.asarray(frame_idx) == f
        z[sel] = (pt[np.asarray(point_idx)[sel]] @ (Rx @ Ry @ Rz).T + vw[f, 3:])[:, 2]
    b = (fL * z) / (z - fL)                                              # :480
    bad = (v < 2) | (b < 0)                                              # :483
    a = np.stack([np.where(bad, 0.0, v), np.where(bad, 0.0, 1.0)], 1)
    x, *_ = np.linalg.lstsq(a, np.where(bad, 0.0, b), rcond=None)       # :492
    return fL, float(x[0]), float(x[1])
