"""BASELINE.json configs[4]: sweep over the number of observations (1e5 .. 1e8), robust vs non-robust cost and 0/1/2
radial distortion parameters, on the GPUs this process group has (1 process = 1 GPU; under torchrun every rank takes its
shard). One JSON line per case on rank 0: M residual+Jacobian evals/s over complete LM solves, LM iterations/s, and the
fused evaluation kernel alone.

  python scripts/sweep_cfg5.py [--max-obs 1e8]
  torchrun --nproc-per-node N scripts/sweep_cfg5.py
Scenes: the cfg4 family (window 4, ~28 observations per (point, frame)); points x frames chosen to hit the target N.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lifcal_b200 import api, capi  # noqa: E402

CASES = {1e5: (900, 40), 1e6: (9000, 100), 1e7: (90000, 300), 1e8: (900000, 1000)}  # N ~ 112 * points (window 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-obs", type=float, default=1e8)
    ap.add_argument("--min-obs", type=float, default=0.0)
    args = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    torch.cuda.set_device(lr)
    comm, dist = None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
        comm = api.Communicator(rank, world, bench.broadcast_unique_id(api.comm_unique_id() if rank == 0 else None))
    for target, (npts, nfr) in CASES.items():
        if target > args.max_obs or target < args.min_obs:
            continue
        for robust in (1, 0):
            for nrad in (0, 1, 2):
                cfg = nrad | capi.CFG_TANGENTIAL | capi.CFG_MLADJ | capi.CFG_REFINE_POSES | capi.CFG_REFINE_POINTS | \
                    (capi.CFG_ROBUST if robust else 0)
                spec = capi.scene_spec(4, n_points=npts, n_frames=nfr, order=1, config=cfg)
                if world > 1:
                    spec.point_begin, spec.point_end = bench.shard_range(npts, rank, world)
                sc = capi.Scene(spec)
                ds = api.DeviceSolver(sc.problem, api.default_options(device=lr), communicator=comm)
                ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
                ds.run()
                s = ds.run()
                ev = ds.time_eval(5, False)
                t = torch.tensor([s["solve_gpu_ms"], ev], dtype=torch.float64, device="cuda")
                if dist is not None:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if rank == 0:
                    n = s["num_observations"]
                    print(json.dumps({"n_gpus": world, "target_obs": target, "observations": n, "robust": robust, "n_radial": nrad,
                                      "lm_rows": s["num_iterations"], "evals": s["num_jacobian_evals"],
                                      "solve_gpu_ms": round(float(t[0]), 3),
                                      "M_evals_s": round(n * s["num_jacobian_evals"] / float(t[0]) / 1e3, 1),
                                      "lm_iters_s": round(s["num_iterations"] / float(t[0]) * 1e3, 1),
                                      "fused_eval_ms_rank_max": round(float(t[1]), 4)}), flush=True)
                ds.close()
    if comm is not None:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
