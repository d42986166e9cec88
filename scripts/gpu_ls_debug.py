"""Debug aid: line-search trace of the device loop and of the oracle on the recalib scene that starts on a bound."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LFBA_DEBUG"] = "1"
from lifcal_b200 import api, capi
from oracle import binding as ob
sc = capi.make_scene(None, n_points=150, n_frames=6, seed=13, calib_type=capi.RECALIBRATION, init_intrinsics_rel=2e-4)
cam0 = sc.camera_init.copy()
cam0[1] = sc.camera_true[1] / 1.3 * 0.9999
init = (cam0, sc.views_init, sc.points_init)
g = api.solve(sc.problem, *init)
sys.stdout.flush()
o = ob.solve(sc.problem, *init)
for r, q in zip(g[3]["iterations"], o[3]["iterations"]):
    print(r["iteration"], r["cost"], q["cost"], r["line_search_iterations"], q["line_search_iterations"], r["step_norm"], q["step_norm"])
