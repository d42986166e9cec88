"""Run on the GPU box: ONE device-resident LM solve of a workload (for ncu launch lists: keep it short)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lifcal_b200 import api, capi
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
kw = {}
if len(sys.argv) > 2:
    kw = dict(n_points=int(sys.argv[2]), n_frames=int(sys.argv[3]))
sc = capi.make_scene(int(name[-1]), order=1, **kw)
ds = api.DeviceSolver(sc.problem)
ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
s = ds.run()
print(name, s["num_observations"], s["num_iterations"], s["solve_gpu_ms"], s["final_cost"])
ds.close()
