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


def make_lens_grid(raw=512, diameter=23.0, rotation=0.003, scale=2, seed=0):
    """A Raytrix-like hexagonal micro-lens grid with the two pixel maps projectPointsToRawImage reads
    (MicroLensGrid::createGrid / defineMlMaps, src/MicroLensGrid/MicroLensGrid.cpp:186-270, 338-421): float32 centres;
    map_ml = lens whose valid micro image (radius D/2 - 1) covers the pixel, map_next = nearest lens. The maps are INPUTS of
    the component under test, so a KD-tree stands in for the reference's ring search."""
    from scipy.spatial import cKDTree
    D = np.float32(diameter)
    hy = np.float32(np.sqrt(0.75)) * D
    ca, sa = np.float32(np.cos(rotation)), np.float32(np.sin(rotation))
    n = int(raw / diameter) + 4
    cs = []
    for j in range(-2, int(raw / float(hy)) + 3):
        for i in range(-2, n):
            x = np.float32(i) * D + (np.float32(0.5) * D if j % 2 else np.float32(0))
            y = np.float32(j) * hy
            cs.append((np.float32(11.3) + x * ca - y * sa, np.float32(7.9) + x * sa + y * ca))
    c = np.array(cs, np.float32)
    yy, xx = np.mgrid[0:raw, 0:raw]
    d, idx = cKDTree(c.astype(np.float64)).query(np.stack([xx.ravel(), yy.ravel()], 1).astype(np.float64))
    g = capi.LensGrid(raw, raw, scale, D, rotation, True, c[:, 0], c[:, 1], idx.astype(np.int32), idx.astype(np.int32))
    g.map_ml = np.where(d * d <= float(g.lens_validity_radius_2), idx, -1).astype(np.int32)
    return g
