"""tests/golden/make_golden.py — regenerates the committed golden vectors.

Run HERE (the container that has /root/reference): the functor outputs are produced by the REFERENCE's own
code — src/BundleAdjustment/BundleAdjustment.h + src/CameraModel.h compiled in place into
oracle/_ref/libref_functor.so (oracle/ref_bridge.cpp) — so the vectors pin the reference's arithmetic and
can travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py

Outputs
  functor_kat.npz   per (model config x arity): inputs, residuals (n,2) and Jet Jacobians (n,2,26) of
                    OurCostFunctionBundle, plus OurConstraintFunctionBundle samples
  solver_small.json Ceres-style iteration tables + final parameters of the oracle LM (reference functor
                    plugged in) on small seeded scenes; "parity unpinned" for the LM loop itself (no Ceres here)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from lifcal_b200 import capi  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ARITIES = {"cam_view_point": capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS, "cam_view": capi.CFG_REFINE_POSES,
           "cam": 0}


def functor_kat(n=24, ncam=2):
    assert ob.ref_lib() is not None, "oracle/_ref/libref_functor.so missing (make -C oracle)"
    rng = np.random.default_rng(20240910)
    out = {}
    for mc in helpers.all_model_configs():
        for aname, abits in ARITIES.items():
            cfg = mc | abits
            for ci in range(ncam):
                b = helpers.random_blocks(rng, cfg, n, signs=(ci == 1))
                cam = b["cams"][0]
                pa = capi.ProblemArrays(cfg, 0, b["spx"], b["spy"], b["scale"], n, n, b["obs"][:, 0], b["obs"][:, 1],
                                        b["ml"][:, 0], b["ml"][:, 1], np.arange(n), np.arange(n))
                ev = ob.evaluate(pa, cam, b["views"].ravel(), b["points"].ravel(), use_ref=True)
                key = f"cfg{cfg:#06x}_{ci}"
                out[key + "_camera"] = cam
                out[key + "_views"] = b["views"]
                out[key + "_points"] = b["points"]
                out[key + "_ml"] = b["ml"]
                out[key + "_obs"] = b["obs"]
                out[key + "_res"] = ev["residuals"]
                out[key + "_jc"] = ev["jac_camera"]
                out[key + "_jv"] = ev["jac_view"]
                out[key + "_jp"] = ev["jac_point"]
    out["spx"] = np.array([b["spx"]])
    out["scale"] = np.array([b["scale"]])
    # distance constraints
    R = ob.ref_lib()
    import ctypes as C
    m = 32
    p1 = 1000 * rng.standard_normal((m, 3))
    p2 = p1 + 200 * rng.standard_normal((m, 3))
    dist = np.linalg.norm(p1 - p2, axis=1) * (1 + 0.01 * rng.standard_normal(m))
    sig = 0.05 + 0.2 * rng.random(m)
    res = np.zeros(m)
    jac = np.zeros((m, 6))
    for i in range(m):
        r = C.c_double()
        j = np.zeros(6)
        R.ref_distance_eval(dist[i], sig[i], capi._dp(np.ascontiguousarray(p1[i])),
                            capi._dp(np.ascontiguousarray(p2[i])), C.byref(r), capi._dp(j))
        res[i] = r.value
        jac[i] = j
    out.update(dc_p1=p1, dc_p2=p2, dc_dist=dist, dc_sigma=sig, dc_res=res, dc_jac=jac)
    np.savez_compressed(os.path.join(HERE, "functor_kat.npz"), **out)
    print("functor_kat.npz:", len(out), "arrays")


SOLVER_CASES = {
    # name: (scene kwargs)
    "tiny_full": dict(n_points=40, n_frames=4, n_constraints=2, seed=11),
    "tiny_nonrobust_rad1": dict(n_points=40, n_frames=4, seed=12,
                                config=1 | capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS | capi.CFG_MLADJ),
    "tiny_recalib": dict(n_points=60, n_frames=5, seed=13, calib_type=capi.RECALIBRATION),
    "tiny_poses_only": dict(n_points=60, n_frames=5, seed=14,
                            config=2 | capi.CFG_TANGENTIAL | capi.CFG_REFINE_POSES | capi.CFG_ROBUST | capi.CFG_MLADJ),
    "tiny_camera_only": dict(n_points=60, n_frames=5, seed=15, config=2 | capi.CFG_TANGENTIAL | capi.CFG_ROBUST),
    "small_window": dict(n_points=300, n_frames=12, window=4, n_constraints=3, seed=16),
}


def solver_small():
    out = {}
    for name, kw in SOLVER_CASES.items():
        sc = capi.make_scene(None, **kw)
        cam, vw, pt, s = ob.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init, use_ref=True, threads=1)
        rows = [{k: r[k] for k in ("iteration", "cost", "cost_change", "gradient_max_norm", "step_norm",
                                    "relative_decrease", "trust_region_radius", "step_is_successful")}
                for r in s["iterations"]]
        out[name] = dict(scene=kw, n_obs=sc.problem.n_obs, num_iterations=s["num_iterations"],
                         stop_reason=s["stop_reason"], termination_type=s["termination_type"],
                         initial_cost=s["initial_cost"], final_cost=s["final_cost"], rows=rows,
                         camera=cam.tolist(), views=vw.tolist(),
                         points_checksum=[float(pt.sum()), float(np.abs(pt).sum())],
                         points_head=pt[:30].tolist())
        print(name, "N", sc.problem.n_obs, "iters", s["num_iterations"], "stop", s["stop_reason"], "cost",
              s["initial_cost"], "->", s["final_cost"])
    with open(os.path.join(HERE, "solver_small.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    functor_kat()
    solver_small()
