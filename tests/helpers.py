"""Shared helpers for the tests (random functor inputs, config enumeration)."""
import numpy as np

from lifcal_b200 import capi


def all_model_configs():
    """(n_radial, tangential, ml_adjust) x flag sets used by the reference (BundleAdjustment.h:28-79)."""
    out = []
    for nrad in (0, 1, 2):
        for tan in (0, 1):
            for mladj in (0, 1):
                out.append(nrad | (capi.CFG_TANGENTIAL if tan else 0) | (capi.CFG_MLADJ if mladj else 0))
    return out


def random_camera(rng, config, signs=False):
    cam = np.zeros(17)
    cam[0] = 35.0 * (1 + 0.01 * rng.standard_normal())
    cam[1] = 33.07 * (1 + 0.005 * rng.standard_normal())
    cam[2] = 0.57 * (1 + 0.02 * rng.standard_normal())
    cam[3] = 511.3 + 3 * rng.standard_normal()
    cam[4] = 512.9 + 3 * rng.standard_normal()
    idx = 5
    nrad = config & 3
    kk = [2e-4, -3e-7]
    for i in range(nrad):
        cam[idx] = kk[i] * (1 + 0.3 * rng.standard_normal())
        idx += 1
    if config & capi.CFG_TANGENTIAL:
        cam[idx] = 1e-5 * (1 + 0.3 * rng.standard_normal())
        cam[idx + 1] = -2e-5 * (1 + 0.3 * rng.standard_normal())
    if signs:  # the functor takes |fL|, |bL0|, |B| (BundleAdjustment.h:123-128)
        for j in range(3):
            if rng.random() < 0.5:
                cam[j] = -cam[j]
    return cam


def random_blocks(rng, config, n, signs=False):
    """n independent residual blocks with their own camera/view/point (problem with n frames, n points)."""
    spx = 2.0 * float(np.float32(0.0055))
    scale = 2.0
    s_raw = spx / scale
    cams = np.stack([random_camera(rng, config, signs) for _ in range(n)])
    views = np.zeros((n, 6))
    views[:, :3] = 0.05 * rng.standard_normal((n, 3))
    views[:, 3:] = 30 * rng.standard_normal((n, 3))
    ml = np.float32(100 + 1848 * rng.random((n, 2))).astype(np.float64)
    # camera-frame point whose virtual image falls near the lens, then moved to world coordinates
    v = 3.0 + 6.0 * rng.random(n)
    bL = 33.07 + v * 0.57
    Z = 35.0 * bL / (bL - 35.0)
    craw = (np.array([511.3, 512.9]) + 0.5) * scale - 0.5
    xv = ml + (rng.random((n, 2)) - 0.5) * 2 * (v[:, None] * 9.0)
    pc = np.stack([(xv[:, 0] - craw[0]) * s_raw * Z / bL, (xv[:, 1] - craw[1]) * s_raw * Z / bL, Z], axis=1)
    pts = np.zeros((n, 3))
    for i in range(n):
        a = views[i, :3]
        cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        R = Rx @ Ry @ Rz
        pts[i] = R.T @ (pc[i] - views[i, 3:])
    obs = np.float32(ml + 8 * (rng.random((n, 2)) - 0.5)).astype(np.float64)
    return dict(spx=spx, spy=spx, scale=scale, cams=cams, views=views, points=pts, ml=ml, obs=obs)
