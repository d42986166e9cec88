"""Run on the GPU box: time the fused evaluation pass (k_eval_rows) and one LM solve per workload.
Environment switches read by the library: LFBA_LANES (lanes per track), LFBA_CHOL_PARTS, LFBA_FRAME_SPLITS."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifcal_b200 import api, capi  # noqa: E402

for name in sys.argv[1:] or ["cfg3"]:
    sc = capi.make_scene(int(name[-1]), order=1)
    ds = api.DeviceSolver(sc.problem)
    ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
    s = ds.run()
    ev = ds.time_eval(10, False)
    n = s["num_observations"]
    print(json.dumps({"lanes": os.environ.get("LFBA_LANES", "auto"), "workload": name, "fused_eval_ms": round(ev, 4),
                      "M_evals_s": round(n / ev / 1e3, 1), "rows": s["num_iterations"], "final_cost": s["final_cost"],
                      "solve_gpu_ms": round(s["solve_gpu_ms"], 3)}), flush=True)
    ds.close()
