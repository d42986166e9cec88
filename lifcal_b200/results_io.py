"""Result files of a LiFCal calibration — the step AFTER the bundle-adjustment hot path.

Writers/readers for the files `CameraCalibration::storeResults()` produces from the solved parameters, in the
reference's exact text formats, so that a harness can diff a result folder written from `lfba_solve()` output against one
written by the reference:

  CameraModel.xml              src/CameraCalibration.cpp:1296-1384  (pugixml, tab indentation, boost::lexical_cast doubles)
  extrinsicOrientations.xml    :1386-1437
  ExtrinsicOrientations.txt    :1439-1482  ("%05d" + 16 x " %16.10f", rows of [R t; 0 1], frames sorted by id)
  rawImagePoints.csv           :1484-1546  ("%d,%d,%f,%f,%f,%f,%d": frame id, i, observed x, y, projected x, y, point index)
  calibrationProtocol.txt      :1548-1617

Host-side Python, no GPU work: the projected coordinates of rawImagePoints.csv are `observed + residual` with the
residuals of `lfba_eval()` (the reference recomputes the projection with the same functor arithmetic, :1523-1530).
`boost::lexical_cast<std::string>(double)` prints 17 significant digits ("%.17g"), which round-trips a double exactly.
"""
from __future__ import annotations

import math
import re
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Sequence

import numpy as np


def _lex(x: float) -> str:
    """boost::lexical_cast<std::string>(double): max_digits10 = 17 significant digits, %g style."""
    return "%.17g" % float(x)


def _xml_lines(tag: str, text: Optional[str] = None, attrs: Optional[Dict[str, str]] = None, children=None, depth=0) -> List[str]:
    """pugixml's default writer (format_indent, one tab per level; an element with only text stays on one line)."""
    ind = "\t" * depth
    a = "".join(f' {k}="{v}"' for k, v in (attrs or {}).items())
    if children:
        out = [f"{ind}<{tag}{a}>"]
        for c in children:
            out += _xml_lines(depth=depth + 1, **c)
        out.append(f"{ind}</{tag}>")
        return out
    if text is None:
        return [f"{ind}<{tag}{a} />"]
    return [f"{ind}<{tag}{a}>{text}</{tag}>"]


def _doc(root_children) -> str:
    return "\n".join(['<?xml version="1.0" encoding="UTF-8"?>'] + _xml_lines("Root", children=root_children)) + "\n"


# ------------------------------------------------------------------------------------------------------------------
# CameraModel.xml
# ------------------------------------------------------------------------------------------------------------------
def camera_model_xml(camera17: Sequence[float], n_radial: int, tangential: bool, ml_center_adjustment: bool,
                     image_size: Sequence[int], pixel_size: float) -> str:
    """camera17 = the 17-wide camera block [fL, bL0, B, cx, cy, k.., t0, t1, 0..] as written back at :967-988."""
    cam = np.asarray(camera17, dtype=np.float64)
    ch = [
        dict(tag="CalibrationModel", text="Plenoptic"),
        dict(tag="ImageSize", attrs={"units": "pix"}, children=[dict(tag="Width", text=str(int(image_size[0]))),
                                                                 dict(tag="Height", text=str(int(image_size[1])))]),
        dict(tag="PixelSize", attrs={"units": "mm"}, text="%.5f" % pixel_size),
        dict(tag="PrincipalPoint", attrs={"units": "pix"}, children=[dict(tag="x", text=_lex(cam[3])), dict(tag="y", text=_lex(cam[4]))]),
        dict(tag="FocalLength", attrs={"units": "mm"}, text=_lex(cam[0])),
        dict(tag="MainLensMlaDistance", attrs={"units": "mm"}, text=_lex(cam[1])),
        dict(tag="SensorMlaDistance", attrs={"units": "mm"}, text=_lex(cam[2])),
    ]
    if n_radial > 0:
        ch.append(dict(tag="RadialDistortion", attrs={"units": "mm"},
                       children=[dict(tag=f"A{i}", text=_lex(cam[5 + i])) for i in range(n_radial)]))
    if tangential:
        ch.append(dict(tag="TangentialDistortion", attrs={"units": "mm"},
                       children=[dict(tag="B0", text=_lex(cam[5 + n_radial])), dict(tag="B1", text=_lex(cam[6 + n_radial]))]))
    ch.append(dict(tag="MicroLensCenterAdjustment", text="true" if ml_center_adjustment else "false"))
    return _doc(ch)


def parse_camera_model_xml(text: str) -> dict:
    root = ET.fromstring(text)
    rad = root.find("RadialDistortion")
    tan = root.find("TangentialDistortion")
    n_radial = len(list(rad)) if rad is not None else 0
    cam = np.zeros(17)
    cam[0] = float(root.find("FocalLength").text)
    cam[1] = float(root.find("MainLensMlaDistance").text)
    cam[2] = float(root.find("SensorMlaDistance").text)
    cam[3] = float(root.find("PrincipalPoint/x").text)
    cam[4] = float(root.find("PrincipalPoint/y").text)
    for i in range(n_radial):
        cam[5 + i] = float(rad.find(f"A{i}").text)
    if tan is not None:
        cam[5 + n_radial] = float(tan.find("B0").text)
        cam[6 + n_radial] = float(tan.find("B1").text)
    return dict(camera=cam, n_radial=n_radial, tangential=tan is not None,
                ml_center_adjustment=root.find("MicroLensCenterAdjustment").text == "true",
                image_size=(int(root.find("ImageSize/Width").text), int(root.find("ImageSize/Height").text)),
                pixel_size=float(root.find("PixelSize").text), model=root.find("CalibrationModel").text)


# ------------------------------------------------------------------------------------------------------------------
# extrinsicOrientations.xml / ExtrinsicOrientations.txt
# ------------------------------------------------------------------------------------------------------------------
def extrinsic_orientations_xml(views6F: Sequence[float], frame_ids: Sequence[int]) -> str:
    v = np.asarray(views6F, dtype=np.float64).reshape(-1, 6)
    ch = []
    for k, fid in enumerate(frame_ids):
        ch.append(dict(tag="Frame", attrs={"id": str(int(fid))}, children=[
            dict(tag="Rotation", children=[dict(tag="Coeff", attrs={"i": str(i)}, text=_lex(v[k, i])) for i in range(3)]),
            dict(tag="Translation", children=[dict(tag="Coeff", attrs={"i": str(i)}, text=_lex(v[k, 3 + i])) for i in range(3)]),
        ]))
    return _doc(ch)


def parse_extrinsic_orientations_xml(text: str):
    root = ET.fromstring(text)
    ids, views = [], []
    for fr in root.findall("Frame"):
        ids.append(int(fr.get("id")))
        rot = {int(c.get("i")): float(c.text) for c in fr.find("Rotation")}
        tr = {int(c.get("i")): float(c.text) for c in fr.find("Translation")}
        views.append([rot[0], rot[1], rot[2], tr[0], tr[1], tr[2]])
    return ids, np.asarray(views, dtype=np.float64).reshape(-1)


def transformation_matrix(view6: Sequence[float]) -> np.ndarray:
    """RigidBody::getTransformationMatrix (src/CameraModel.h:246-264): [Rx(a0) Ry(a1) Rz(a2) | t; 0 0 0 1]."""
    a0, a1, a2 = view6[0], view6[1], view6[2]
    c0, s0, c1, s1, c2, s2 = math.cos(a0), math.sin(a0), math.cos(a1), math.sin(a1), math.cos(a2), math.sin(a2)
    Rx = np.array([[1, 0, 0], [0, c0, -s0], [0, s0, c0]])
    Ry = np.array([[c1, 0, s1], [0, 1, 0], [-s1, 0, c1]])
    Rz = np.array([[c2, -s2, 0], [s2, c2, 0], [0, 0, 1]])
    M = np.eye(4)
    M[:3, :3] = Rx @ Ry @ Rz
    M[:3, 3] = view6[3:6]
    return M


def extrinsic_orientations_txt(views6F: Sequence[float], frame_ids: Sequence[int]) -> str:
    v = np.asarray(views6F, dtype=np.float64).reshape(-1, 6)
    order = sorted(range(len(frame_ids)), key=lambda k: frame_ids[k])  # :1450-1456 (std::sort by id)
    lines = []
    for k in order:
        M = transformation_matrix(v[k])
        lines.append("%05d" % int(frame_ids[k]) + "".join(" %16.10f" % M[y, x] for y in range(4) for x in range(4)))
    return "\n".join(lines) + ("\n" if lines else "")


def parse_extrinsic_orientations_txt(text: str):
    ids, mats = [], []
    for ln in text.splitlines():
        tok = ln.split()
        if not tok:
            continue
        ids.append(int(tok[0]))
        mats.append(np.asarray([float(t) for t in tok[1:17]]).reshape(4, 4))
    return ids, mats


# ------------------------------------------------------------------------------------------------------------------
# rawImagePoints.csv
# ------------------------------------------------------------------------------------------------------------------
def raw_image_points_csv(obs_x, obs_y, residuals, point_idx, frame_idx, frame_ids: Sequence[int]) -> str:
    """Observations in the reference's frame-major order (:1501-1541); `residuals` [N,2] from lfba_eval():
    projected = observed + residual. The per-frame running index i restarts at 0 for every frame."""
    res = np.asarray(residuals, dtype=np.float64).reshape(-1, 2)
    frame_idx = np.asarray(frame_idx)
    out = []
    counter: Dict[int, int] = {}
    last = None
    for k in range(len(obs_x)):
        f = int(frame_idx[k])
        if last is not None and f < last:
            raise ValueError("rawImagePoints.csv needs the frame-major observation order of the reference")
        last = f
        i = counter.get(f, 0)
        counter[f] = i + 1
        out.append("%d,%d,%f,%f,%f,%f,%d" % (int(frame_ids[f]), i, obs_x[k], obs_y[k], obs_x[k] + res[k, 0], obs_y[k] + res[k, 1],
                                              int(point_idx[k])))
    return "\n".join(out) + ("\n" if out else "")


def parse_raw_image_points_csv(text: str) -> np.ndarray:
    rows = [ln.split(",") for ln in text.splitlines() if ln]
    return np.asarray([[float(t) for t in r] for r in rows], dtype=np.float64).reshape(-1, 7)


# ------------------------------------------------------------------------------------------------------------------
# calibrationProtocol.txt
# ------------------------------------------------------------------------------------------------------------------
def calibration_protocol(camera17, n_radial: int, tangential: bool, ml_center_adjustment: bool, pixel_size: float,
                         refine_poses: bool, refine_points: bool, robust: bool, stats: dict) -> str:
    """stats: std_x, std_y, mae_x, mae_y as returned by lfba_eval() (calcReprojectionError, :1026-1103)."""
    cam = np.asarray(camera17, dtype=np.float64)
    s = ("*******************************************************************************\n"
         "***   LiFCal: Online Light Field Camera Calibration via Bundle Adjustment   ***\n"
         "*******************************************************************************\n\n")
    s += "*** Intrinsic Parameters ***\n"
    s += "Pixel Size: %1.3f mm\n" % pixel_size
    for name, val in (("fL  ", cam[0]), ("bL0 ", cam[1]), ("B   ", cam[2]), ("cx  ", cam[3]), ("cy  ", cam[4])):
        s += "\t%s : %18.15f\n" % (name, val)
    for i in range(n_radial):
        s += "\ta%d   : %18.15f\n" % (i, cam[5 + i])
    if tangential:
        s += "\tb0   : %18.15f\n" % cam[5 + n_radial]
        s += "\tb1   : %18.15f\n" % cam[6 + n_radial]
    s += "\n"
    if ml_center_adjustment:
        s += "\tDid micro lens center adjustment\n"
    s += "*** Additional Settings ***\n\tDistortion defined on MLA plane.\n\n"
    s += ("\tExtrinsic Orientations were refined.\n" if refine_poses else "\tExtrinsic Orientations from COLMAP were kept.\n") + "\n"
    s += ("\t3D Object coordinates were refined.\n" if refine_points else "\t3D Object coordinates from COLMAP were kept.\n") + "\n"
    s += ("\tRobust cost function was used for estimation.\n" if robust else "\tSquared cost function was used for estimation.\n") + "\n"
    s += "*** Statistics ***\n\tReprojection errors:\n"
    s += "\tstd. Dev. x:           %8.5f\n" % stats["std_x"]
    s += "\tstd. Dev. y:           %8.5f\n" % stats["std_y"]
    s += "\tmae x:                 %8.5f\n" % stats["mae_x"]
    s += "\tmae y:                 %8.5f\n" % stats["mae_y"]
    return s


_PROTO_RE = re.compile(r"^\t(fL|bL0|B|cx|cy|a\d|b\d)\s*:\s*(-?\d+\.\d+)$", re.M)


def parse_calibration_protocol(text: str) -> Dict[str, float]:
    return {m.group(1): float(m.group(2)) for m in _PROTO_RE.finditer(text)}
