"""torchrun target: N-rank sharded solve vs the CPU oracle on the same scene (no constraints), rank 0 reports."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import bench
from lifcal_b200 import api, capi
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
uid = bench.broadcast_unique_id(api.comm_unique_id() if rank == 0 else None)
comm = api.Communicator(rank, world, uid)
for kw in [dict(n_points=2000, n_frames=24, window=4, seed=5), dict(n_points=600, n_frames=8, seed=6)]:
    P = kw["n_points"]
    lo, hi = bench.shard_range(P, rank, world)
    sc = capi.make_scene(None, order=1, point_begin=lo, point_end=hi, **kw)
    ds = api.DeviceSolver(sc.problem, api.default_options(device=lr), communicator=comm)
    ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
    s = ds.run()
    cam, vw, pt = ds.get_parameters()
    # gather the owned points on rank 0
    t = torch.from_numpy(pt.copy()).cuda()
    mask = torch.zeros(3 * P, dtype=torch.float64, device="cuda")
    mask[3 * lo:3 * hi] = 1
    t = t * mask
    dist.all_reduce(t)
    if rank == 0:
        from oracle import binding as ob
        full = capi.make_scene(None, order=1, **kw)
        o1 = ob.solve(full.problem, full.camera_init, full.views_init, full.points_init, threads=1)
        oN = ob.solve(full.problem, full.camera_init, full.views_init, full.points_init, threads=8)
        ptg = t.cpu().numpy()
        un = np.abs(ptg) == 0
        ptg[un] = full.points_init[un]
        crel = max(abs(a["cost"] - b["cost"]) / b["cost"] for a, b in zip(s["iterations"], oN[3]["iterations"]))
        live = np.abs(oN[0]) > 0
        print(f"world={world} {kw}: rows {s['num_iterations']}/{oN[3]['num_iterations']} cost_rel_max {crel:.2e} "
              f"cam rel max {np.max(np.abs(cam - oN[0])[live] / np.abs(oN[0])[live]):.2e} (oracle spread {np.max(np.abs(o1[0] - oN[0])[live] / np.abs(oN[0])[live]):.2e}) "
              f"views {np.max(np.abs(vw - oN[1])):.2e} points {np.max(np.abs(ptg - oN[2])):.2e} n_obs {s['num_observations']}/{full.problem.n_obs}", flush=True)
    ds.close()
comm.close()
dist.destroy_process_group()
