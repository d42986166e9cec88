"""Initial plenoptic parameters — the step BEFORE the bundle-adjustment hot path.

`CameraCalibration::initPlenopticParameters()` (src/CameraCalibration.cpp:456-498): fL_init = fPH_init * pixelSize_totFoc,
then the linear model  bL = v * B + bL0  is fitted over all (frame, point) pairs, with  bL = fL Z / (Z - fL)  from the
camera-frame depth Z of the COLMAP point and v the measured virtual depth; rows with v < 2 or bL < 0 are zeroed (they
stay in the system as all-zero rows, which leaves the least-squares solution unchanged).
`initPlenopticParametersRecalibration()` (:503-514): bL0_init = fL_init - 2 B_init from the fixed parameters.

Host-side (two unknowns, five sums): numpy. Inputs are flat arrays over all (frame, point) pairs.
"""
from __future__ import annotations

import numpy as np


def init_plenoptic_parameters(fph_init: float, pixel_size_totfoc: float, virtual_depth, z_cam):
    """Returns (fL_init, B_init, bL0_init). virtual_depth[k], z_cam[k]: virtual depth and camera-frame Z of pair k."""
    fL = float(fph_init) * float(pixel_size_totfoc)
    v = np.asarray(virtual_depth, dtype=np.float64)
    z = np.asarray(z_cam, dtype=np.float64)
    b = fL * z / (z - fL)
    bad = (v < 2.0) | (b < 0.0)          # :483-488
    a0 = np.where(bad, 0.0, v)
    a1 = np.where(bad, 0.0, 1.0)
    bb = np.where(bad, 0.0, b)
    # normal equations of [a0 a1] x = bb (the reference uses a thin SVD; same minimiser for a full-rank 2-column system)
    s00, s01, s11 = float(a0 @ a0), float(a0 @ a1), float(a1 @ a1)
    r0, r1 = float(a0 @ bb), float(a1 @ bb)
    det = s00 * s11 - s01 * s01
    if not det > 0.0:
        raise ValueError("degenerate initialisation: fewer than two distinct valid virtual depths")
    B = (s11 * r0 - s01 * r1) / det
    bL0 = (s00 * r1 - s01 * r0) / det
    return fL, B, bL0


def init_plenoptic_parameters_recalibration(fL_fixed: float, B_fixed: float):
    """(:503-514) returns (fL_init, B_init, bL0_init)."""
    return float(fL_fixed), float(B_fixed), float(fL_fixed) - 2.0 * float(B_fixed)
