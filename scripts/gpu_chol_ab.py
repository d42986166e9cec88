"""Run on the GPU box: the partitioned reduced solve against the single-chain banded one on the same scene
(LFBA_CHOL_PARTS=1 disables the partitioned path; the variable is read when a solver is created)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from lifcal_b200 import api, capi  # noqa: E402

def run(sc, parts):
    if parts is None:
        os.environ.pop("LFBA_CHOL_PARTS", None)
    else:
        os.environ["LFBA_CHOL_PARTS"] = str(parts)
    ds = api.DeviceSolver(sc.problem, api.default_options(profile=1))
    ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
    ds.run()
    s = ds.run()
    p = ds.get_parameters()
    ds.close()
    return s, p

for name in sys.argv[1:] or ["mid"]:
    sc = capi.make_scene(int(name[-1]), order=1) if name.startswith("cfg") else capi.make_scene(None, n_points=1500, n_frames=64, window=4, seed=77, order=1)
    s0, p0 = run(sc, 1)
    for parts in [None, 8, 24, 32]:
        s1, p1 = run(sc, parts)
        dc = max(abs(a["cost"] - b["cost"]) / b["cost"] for a, b in zip(s1["iterations"], s0["iterations"]))
        print(json.dumps({"scene": name, "parts": parts, "rows": [s0["num_iterations"], s1["num_iterations"]], "cost_rel_max": dc,
                          "cam_rel": float(np.max(np.abs(p1[0][:9] - p0[0][:9]) / np.abs(p0[0][:9]))),
                          "views_abs": float(np.max(np.abs(p1[1] - p0[1]))),
                          "chol_ms": [round(s0["kernel_ms"]["cholesky"] / s0["kernel_calls"]["cholesky"], 4),
                                      round(s1["kernel_ms"]["cholesky"] / s1["kernel_calls"]["cholesky"], 4)],
                          "solve_ms": [round(s0["solve_gpu_ms"], 2), round(s1["solve_gpu_ms"], 2)]}), flush=True)
