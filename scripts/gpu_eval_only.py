"""Run on the GPU box: time the eval-only (Jacobian materialised) kernel."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifcal_b200 import api, capi  # noqa: E402
for name in sys.argv[1:] or ["cfg3"]:
    sc = capi.make_scene(int(name[-1]), order=1)
    ds = api.DeviceSolver(sc.problem)
    ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
    n = sc.problem.n_obs
    ms = ds.time_eval(5, True)
    print(json.dumps({"workload": name, "eval_only_ms": round(ms, 4), "GBs_460B": round(460 * n / ms / 1e6, 1), "M_evals_s": round(n / ms / 1e3, 1)}), flush=True)
    ds.close()
