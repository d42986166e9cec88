"""Run on the GPU box: end-to-end lfba_solve() from pinned host buffers (H2D + indexing + LM solve + D2H), repeated."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (pinned pages only)
from lifcal_b200 import api, capi  # noqa: E402
import bench  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = capi.make_scene(int(name[-1]), order=order)
ppa = bench.pinned_problem(sc.problem)
for k in range(4):
    torch.cuda.synchronize()
    t = time.perf_counter()
    cam, vw, pt, s = api.solve(ppa, sc.camera_init, sc.views_init, sc.points_init)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(json.dumps({"workload": name, "order": order, "e2e_s": round(dt, 4), "setup_s": round(s["setup_time_s"], 4),
                      "solve_gpu_ms": round(s["solve_gpu_ms"], 2), "h2d_ms": round(s["h2d_ms"], 2),
                      "h2d_gbs": s["h2d_gbs"] and round(s["h2d_gbs"], 1), "rows": s["num_iterations"], "final_cost": s["final_cost"]}), flush=True)
