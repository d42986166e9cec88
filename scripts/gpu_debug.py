"""Run on the GPU box: GPU solve vs oracle, side-by-side iteration tables (debug aid, not a test)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lifcal_b200 import api, capi  # noqa: E402
from oracle import binding as ob  # noqa: E402

CASES = {
    "tiny_full": dict(n_points=40, n_frames=4, n_constraints=2, seed=11),
    "tiny_nonrobust_rad1": dict(n_points=40, n_frames=4, seed=12,
                                config=1 | capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS | capi.CFG_MLADJ),
    "tiny_poses_only": dict(n_points=60, n_frames=5, seed=14,
                            config=2 | capi.CFG_TANGENTIAL | capi.CFG_REFINE_POSES | capi.CFG_ROBUST | capi.CFG_MLADJ),
    "tiny_camera_only": dict(n_points=60, n_frames=5, seed=15, config=2 | capi.CFG_TANGENTIAL | capi.CFG_ROBUST),
    "small_window": dict(n_points=300, n_frames=12, window=4, n_constraints=3, seed=16),
    "tiny_recalib": dict(n_points=60, n_frames=5, seed=13, calib_type=capi.RECALIBRATION),
    "cfg1": dict(preset=1),
}


def main():
    names = sys.argv[1:] or list(CASES)
    print("devices:", api.device_count())
    for name in names:
        kw = dict(CASES[name])
        preset = kw.pop("preset", None)
        sc = capi.make_scene(preset, **kw)
        t = time.time()
        try:
            ev = api.evaluate(sc.problem, sc.camera_init, sc.views_init, sc.points_init)
            oe = ob.evaluate(sc.problem, sc.camera_init, sc.views_init, sc.points_init)
            print(f"[{name}] eval: max|dr|={np.max(np.abs(ev['residuals'] - oe['residuals'])):.3e} "
                  f"cost gpu={ev['cost']:.12e} oracle={oe['cost']:.12e}")
            for k in ("jac_camera", "jac_view", "jac_point"):
                den = max(1e-300, np.max(np.abs(oe[k])))
                print(f"    {k}: max abs diff / max = {np.max(np.abs(ev[k] - oe[k])) / den:.3e}")
        except Exception as e:  # noqa: BLE001
            print(f"[{name}] eval FAILED: {e}")
        try:
            cam, vw, pt, s = api.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init, raise_on_failure=False)
        except Exception as e:  # noqa: BLE001
            print(f"[{name}] solve FAILED: {e}")
            continue
        tg = time.time() - t
        ocam, ovw, opt_, os_ = ob.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init)
        print(f"[{name}] N={sc.problem.n_obs} status={s['status']} gpu rows={s['num_iterations']} stop={s['stop_reason']} "
              f"oracle rows={os_['num_iterations']} stop={os_['stop_reason']} n={s['reduced_system_size']} "
              f"tracks={s['num_tracks']} lenses={s['num_lenses']} launches={s['gpu_launches']} gpu_ms={s['solve_gpu_ms']:.2f} "
              f"setup_s={s['setup_time_s']:.3f} wall={tg:.2f}")
        for i in range(max(len(s["iterations"]), len(os_["iterations"]))):
            g = s["iterations"][i] if i < len(s["iterations"]) else None
            o = os_["iterations"][i] if i < len(os_["iterations"]) else None
            fmt = lambda r: ("%3d %d cost=%.12e dc=%.3e |g|=%.3e |s|=%.3e rho=%.6f mu=%.3e" % (
                r["iteration"], r["step_is_successful"], r["cost"], r["cost_change"], r["gradient_max_norm"],
                r["step_norm"], r["relative_decrease"], r["trust_region_radius"])) if r else "-"
            print("   G " + fmt(g))
            print("   O " + fmt(o))
        if s["num_iterations"] > 0:
            print(f"   final cost gpu={s['final_cost']:.12e} oracle={os_['final_cost']:.12e} "
                  f"rel={abs(s['final_cost'] - os_['final_cost']) / os_['final_cost']:.3e}")
            print(f"   camera rel diff max = {np.max(np.abs(cam[:9] - ocam[:9]) / (np.abs(ocam[:9]) + 1e-300)):.3e}; "
                  f"views abs diff max = {np.max(np.abs(vw - ovw)):.3e}; points abs diff max = {np.max(np.abs(pt - opt_)):.3e}")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
