"""Run on the GPU box: time the fused evaluation pass for one LFBA_EVAL_MODE (env, read once per process)."""
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
    print(json.dumps({"mode": os.environ.get("LFBA_EVAL_MODE", "default"), "workload": name, "fused_eval_ms": round(ev, 4),
                      "M_evals_s": round(n / ev / 1e3, 1), "rows": s["num_iterations"], "final_cost": s["final_cost"],
                      "solve_gpu_ms": round(s["solve_gpu_ms"], 3)}), flush=True)
    ds.close()
