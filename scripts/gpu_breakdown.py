"""Run on the GPU box: per-stage CUDA-event breakdown of the LM loop for a workload (opt.profile=1)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifcal_b200 import api, capi  # noqa: E402


def main():
    for name in sys.argv[1:] or ["cfg2", "cfg3"]:
        preset = int(name[-1])
        t = time.time()
        sc = capi.make_scene(preset, order=1)
        tg = time.time() - t
        o = api.default_options(profile=1)
        t = time.time()
        ds = api.DeviceSolver(sc.problem, o)
        ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
        ts = time.time() - t
        ds.run()
        s = ds.run()
        n = s["num_observations"]
        print(json.dumps({"workload": name, "N": n, "tracks": s["num_tracks"], "lenses": s["num_lenses"],
                          "n_red": s["reduced_system_size"], "rows": s["num_iterations"], "evals": s["num_jacobian_evals"],
                          "gen_s": round(tg, 2), "setup_s": round(ts, 3), "solve_gpu_ms": round(s["solve_gpu_ms"], 3),
                          "ms_per_iter": round(s["solve_gpu_ms"] / max(1, s["num_iterations"]), 3),
                          "final_cost": s["final_cost"], "launches": s["gpu_launches"],
                          "kernel_ms_per_round": {k: round(v / s["kernel_calls"][k], 4) for k, v in s["kernel_ms"].items()}}))
        o2 = api.default_options()
        ds2 = api.DeviceSolver(sc.problem, o2)
        ds2.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
        ds2.run()
        s2 = ds2.run()
        ev = ds2.time_eval(10, False)
        evm = ds2.time_eval(3, 1)
        evc = ds2.time_eval(3, 2)
        print(json.dumps({"workload": name, "unprofiled_solve_gpu_ms": round(s2["solve_gpu_ms"], 3),
                          "solve_wall_s": round(s2["solve_time_s"], 4), "fused_eval_ms": round(ev, 4),
                          "fused_M_evals_s": round(n / ev / 1e3, 1), "fused_GBs": round((20 * n + 288 * s["num_tracks"]) / ev / 1e6, 1),
                          "eval_only_ceres_layout_ms": round(evm, 4), "eval_only_ceres_GBs_460B": round(460 * n / evm / 1e6, 1),
                          "eval_only_compact_ms": round(evc, 4), "eval_only_compact_GBs_344B": round(344 * n / evc / 1e6, 1),
                          "eval_only_compact_M_evals_s": round(n / evc / 1e3, 1)}))
        ds.close()
        ds2.close()
        sys.stdout.flush()
    print("fp64 peak TFLOP/s:", api.measure_fp64_peak())


if __name__ == "__main__":
    main()
