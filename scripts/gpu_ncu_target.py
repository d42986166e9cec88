"""Small target for ncu: one device-resident LM solve of a workload, capped at a few iterations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifcal_b200 import api, capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = capi.make_scene(int(name[-1]), order=1)
ds = api.DeviceSolver(sc.problem, api.default_options(max_num_iterations=iters))
ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
s = ds.run()
print(name, "rows", s["num_iterations"], "cost", s["final_cost"], "gpu_ms", s["solve_gpu_ms"], "launches", s["gpu_launches"])
if len(sys.argv) > 3:
    print("eval-only ms (Ceres layout)", ds.time_eval(1, 1))
    print("eval-only ms (live camera columns)", ds.time_eval(1, 2))
ds.close()
