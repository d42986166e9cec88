"""CPU tests (no GPU) of lifcal_b200/results_io.py — the result files LiFCal writes after the bundle adjustment
(reference: src/CameraCalibration.cpp:1296-1617). Round trips and the exact text layout."""
import re

import numpy as np
from scipy.spatial.transform import Rotation

from lifcal_b200 import capi, results_io as rio
from oracle import binding as ob


def _solved_scene():
    sc = capi.make_scene(None, n_points=60, n_frames=5, seed=3, order=0)  # order 0 = the reference's frame-major order
    cam, vw, pt, _ = ob.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init, threads=2)
    return sc, cam, vw, pt


def test_camera_model_xml_round_trip_and_layout():
    rng = np.random.default_rng(0)
    for n_radial in (0, 1, 2):
        for tan in (False, True):
            cam = np.zeros(17)
            cam[:5] = [35.0 + rng.standard_normal(), 33.07, 0.57, 511.3, 512.9]
            cam[5:5 + n_radial + 2 * tan] = 1e-4 * rng.standard_normal(n_radial + 2 * tan)
            txt = rio.camera_model_xml(cam, n_radial, tan, True, (1024, 1024), 0.0055)
            back = rio.parse_camera_model_xml(txt)
            assert np.array_equal(back["camera"], cam)  # 17 significant digits: exact
            assert back["n_radial"] == n_radial and back["tangential"] == tan and back["ml_center_adjustment"] is True
            assert back["image_size"] == (1024, 1024) and back["pixel_size"] == 0.0055 and back["model"] == "Plenoptic"
            lines = txt.splitlines()
            assert lines[0] == '<?xml version="1.0" encoding="UTF-8"?>' and lines[1] == "<Root>" and lines[-1] == "</Root>"
            assert lines[2] == "\t<CalibrationModel>Plenoptic</CalibrationModel>"
            assert "\t<PixelSize units=\"mm\">0.00550</PixelSize>" in lines
            assert ("\t<RadialDistortion units=\"mm\">" in lines) == (n_radial > 0)
            assert ("\t\t<B1>" in txt) == tan
    assert rio._lex(35.0) == "35" and rio._lex(0.57) == "0.56999999999999995"  # boost::lexical_cast<std::string>(double)


def test_extrinsic_orientations_xml_and_txt():
    sc, cam, vw, pt = _solved_scene()
    ids = [7, 3, 11, 0, 5]
    txt = rio.extrinsic_orientations_xml(vw, ids)
    bid, bv = rio.parse_extrinsic_orientations_xml(txt)
    assert bid == ids and np.array_equal(bv, vw)
    assert '\t<Frame id="7">' in txt and '\t\t\t<Coeff i="2">' in txt
    t = rio.extrinsic_orientations_txt(vw, ids)
    tid, mats = rio.parse_extrinsic_orientations_txt(t)
    assert tid == sorted(ids)  # the reference sorts the frames by id (:1450-1456)
    v = vw.reshape(-1, 6)
    for fid, M in zip(tid, mats):
        k = ids.index(fid)
        R = Rotation.from_euler("XYZ", v[k, :3]).as_matrix()  # intrinsic X-Y-Z = Rx Ry Rz (src/CameraModel.h:251-254)
        assert np.allclose(M[:3, :3], R, atol=1e-10) and np.allclose(M[:3, 3], v[k, 3:], atol=1e-10)
        assert np.array_equal(M[3], [0, 0, 0, 1])
    first = t.splitlines()[0]
    assert re.fullmatch(r"\d{5}( +-?\d+\.\d{10}){16}", first) and len(first) == 5 + 16 * 17


def test_raw_image_points_csv_and_protocol():
    sc, cam, vw, pt = _solved_scene()
    pa = sc.problem
    ev = ob.evaluate(pa, cam, vw, pt, jacobians=False)
    ids = list(range(100, 100 + pa.n_frames))
    txt = rio.raw_image_points_csv(pa.obs_x, pa.obs_y, ev["residuals"], pa.point_idx, pa.frame_idx, ids)
    tab = rio.parse_raw_image_points_csv(txt)
    assert tab.shape == (pa.n_obs, 7)
    assert np.array_equal(tab[:, 0], 100 + pa.frame_idx) and np.array_equal(tab[:, 6], pa.point_idx)
    for f in range(pa.n_frames):  # the running index restarts in every frame
        sel = tab[pa.frame_idx == f, 1]
        assert np.array_equal(sel, np.arange(len(sel)))
    r = ev["residuals"].reshape(-1, 2)
    assert np.allclose(tab[:, 4] - tab[:, 2], r[:, 0], atol=2e-6) and np.allclose(tab[:, 5] - tab[:, 3], r[:, 1], atol=2e-6)
    assert re.fullmatch(r"\d+,\d+,-?\d+\.\d{6},-?\d+\.\d{6},-?\d+\.\d{6},-?\d+\.\d{6},\d+", txt.splitlines()[0])
    proto = rio.calibration_protocol(cam, 2, True, True, 0.0055, True, True, True, ev["stats"])
    vals = rio.parse_calibration_protocol(proto)
    assert abs(vals["fL"] - cam[0]) < 1e-14 * 40 and abs(vals["a1"] - cam[6]) < 1e-15 and abs(vals["b1"] - cam[8]) < 1e-15
    assert "\tRobust cost function was used for estimation.\n" in proto and "Pixel Size: 0.005 mm\n" in proto or "Pixel Size: 0.006 mm\n" in proto
    assert "\tstd. Dev. x:           %8.5f\n" % ev["stats"]["std_x"] in proto


def test_recalibration_initialisation():
    # src/CameraCalibration.cpp:503-514 (host arithmetic in the reference too); the least-squares fit of :456-498 runs on the
    # device and is tested in tests/test_gpu_project_raw.py
    from lifcal_b200 import init_params as ip
    assert ip.init_plenoptic_parameters_recalibration(35.0, 0.57) == (35.0, 0.57, 35.0 - 1.14)
