"""Run on the GPU box: worst parameter deviation GPU vs oracle, in units of the 1e-9 relative tolerance and of the
oracle's own 1-thread-vs-N-thread spread, for a set of scenes."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifcal_b200 import api, capi
from oracle import binding as ob
cases = json.load(open(os.path.join(ROOT, "tests", "golden", "solver_small.json")))
scenes = [(k, dict(v["scene"])) for k, v in cases.items()] + [("cfg1", dict(preset=1))] + \
         [(f"var{n}", dict(n_points=200, n_frames=5, seed=100 + n, config=n | 0x4 | 0x800 | 0x100 | 0x400 | 0x200)) for n in (0, 1, 2)]
for name, kw in scenes:
    preset = kw.pop("preset", None)
    sc = capi.make_scene(preset, **kw)
    init = (sc.camera_init, sc.views_init, sc.points_init)
    cam, vw, pt, s = api.solve(sc.problem, *init)
    o1 = ob.solve(sc.problem, *init, threads=1)
    oN = ob.solve(sc.problem, *init, threads=max(2, ob.max_threads()))
    live = np.abs(oN[0]) > 0
    dg = np.abs(cam - oN[0])[live] / np.abs(oN[0])[live]
    do = np.abs(o1[0] - oN[0])[live] / np.abs(oN[0])[live]
    crel = max(abs(a["cost"] - b["cost"]) / b["cost"] for a, b in zip(s["iterations"], oN[3]["iterations"]))
    print(f"{name:22s} iters {s['num_iterations']:3d}/{oN[3]['num_iterations']:3d} cost_rel_max {crel:.1e}  cam rel: gpu-vs-oracle max {dg.max():.2e}  "
          f"oracle(1thr)-vs-oracle(Nthr) max {do.max():.2e}  points abs gpu {np.max(np.abs(pt - oN[2])):.1e} oracle-spread {np.max(np.abs(o1[2] - oN[2])):.1e}")
