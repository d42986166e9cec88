"""Run on the GPU box: per-stage breakdown of the LM round for several values of an environment knob that the library
reads when a solver handle is created. Usage: gpu_stage_sweep.py <workload> <ENV_NAME> v1 v2 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifcal_b200 import api, capi  # noqa: E402

name, env, vals = sys.argv[1], sys.argv[2], sys.argv[3:]
sc = capi.make_scene(int(name[-1]), order=1)
for v in vals:
    os.environ[env] = v
    ds = api.DeviceSolver(sc.problem, api.default_options(profile=1))
    ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
    ds.run()
    s = ds.run()
    print(f"{name} {env}={v}: solve {s['solve_gpu_ms']:.3f} ms", {k: round(x / s["kernel_calls"][k], 4) for k, x in s["kernel_ms"].items()}, flush=True)
    ds.close()
