"""Initial plenoptic parameters — the step BEFORE the bundle-adjustment hot path.

`CameraCalibration::initPlenopticParameters()` (src/CameraCalibration.cpp:456-498): fL_init = fPH_init * pixelSize_totFoc,
then the linear model  bL = v * B + bL0  is fitted over all (frame, feature) pairs, with  bL = fL Z / (Z - fL)  from the
camera-frame depth Z of the feature's 3-D point and v the measured virtual depth; rows with v < 2 or bL < 0 are zeroed.
The fit runs on the device (lfba_init_plenoptic in liblfba.so: two fixed-order reduction passes); this module is the
host-side mirror of the two reference entry points. No CPU fallback.
`initPlenopticParametersRecalibration()` (:503-514) is plain host arithmetic in the reference too.
"""
from __future__ import annotations

from . import api


def init_plenoptic_parameters(fph_init, pixel_size_totfoc, virtual_depth, frame_idx, point_idx, views, points):
    """Returns (fL_init, B_init, bL0_init). Pair k: virtual depth virtual_depth[k] of point point_idx[k] seen in frame
    frame_idx[k]; views[6F] = Euler angles + translation of worldToCam, points[3P] world coordinates."""
    return api.init_plenoptic(fph_init, pixel_size_totfoc, virtual_depth, frame_idx, point_idx, views, points)


def init_plenoptic_parameters_recalibration(fL_fixed: float, B_fixed: float):
    """(:503-514) returns (fL_init, B_init, bL0_init)."""
    return float(fL_fixed), float(B_fixed), float(fL_fixed) - 2.0 * float(B_fixed)
