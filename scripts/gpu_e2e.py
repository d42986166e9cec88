"""Run on the GPU box: time the drop-in call lfba_solve() with pinned host buffers (setup phases with LFBA_DEBUG=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lifcal_b200 import api, capi
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
sc = capi.make_scene(int(name[-1]), order=1)
ppa = bench.pinned_problem(sc.problem)
for k in range(3):
    t = time.perf_counter()
    cam, vw, pt, s = api.solve(ppa, sc.camera_init, sc.views_init, sc.points_init)
    dt = time.perf_counter() - t
    print(f"{name} e2e solve {k}: {dt:.3f} s (setup {s['setup_time_s']:.3f} s, solve {s['solve_time_s']:.3f} s, gpu {s['solve_gpu_ms']:.1f} ms) "
          f"N={s['num_observations']} evals={s['num_jacobian_evals']} -> {s['num_observations'] * s['num_jacobian_evals'] / dt / 1e6:.0f} M evals/s", flush=True)
