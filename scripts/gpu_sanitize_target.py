"""Small target for compute-sanitizer: complete LM solves of small scenes through the C ABI with ONE lane per track forced
(LFBA_LANES=1: the large-problem instantiations of the fused kernel), a windowed scene (partitioned reduced solve), a
recalibration scene (projected line search) and the eval-only path."""
import os
import sys

os.environ.setdefault("LFBA_LANES", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifcal_b200 import api, capi  # noqa: E402

cases = [dict(n_points=400, n_frames=6, seed=3, n_constraints=2),
         dict(n_points=1500, n_frames=48, seed=5, window=4),
         dict(n_points=300, n_frames=5, seed=9, calib_type=capi.RECALIBRATION)]
for kw in cases:
    try:
        sc = capi.make_scene(None, **kw)
    except AttributeError as e:  # a keyword this scene generator does not know
        print("skipped", kw, e)
        continue
    cam, vw, pt, s = api.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init)
    ev = api.evaluate(sc.problem, cam, vw, pt)
    print(kw, "N", sc.problem.n_obs, "rows", s["num_iterations"], "cost", s["final_cost"], "eval cost", ev["cost"], flush=True)
print("done")
