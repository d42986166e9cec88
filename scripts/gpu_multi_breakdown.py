"""torchrun target: per-stage CUDA-event breakdown of the sharded LM round (opt.profile = 1) on every rank; rank 0 prints
the max over ranks of each stage (the all-reduce slots include waiting for the slowest rank)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from lifcal_b200 import api, capi
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
uid = bench.broadcast_unique_id(api.comm_unique_id() if rank == 0 else None)
comm = api.Communicator(rank, world, uid)
spec = capi.scene_spec(4, order=1)
lo, hi = bench.shard_range(spec.n_points, rank, world)
spec.point_begin, spec.point_end = lo, hi
sc = capi.Scene(spec)
for prof in (1, 0):
    ds = api.DeviceSolver(sc.problem, api.default_options(device=lr, profile=prof), communicator=comm)
    ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
    ds.run()
    dist.barrier()
    s = ds.run()
    keys = list(s["kernel_ms"].keys())
    per = [s["kernel_ms"][k] / max(1, s["kernel_calls"][k]) for k in keys] + [s["solve_gpu_ms"], float(sc.problem.n_obs), float(s["num_tracks"])]
    t = torch.tensor(per, dtype=torch.float64, device="cuda")
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "profile": prof, "rows": s["num_iterations"],
                          "max_over_ranks": {k: round(float(v), 4) for k, v in zip(keys + ["solve_gpu_ms", "n_obs", "tracks"], tmax.tolist())},
                          "min_over_ranks": {k: round(float(v), 4) for k, v in zip(keys + ["solve_gpu_ms", "n_obs", "tracks"], tmin.tolist())}}), flush=True)
    ds.close()
comm.close()
dist.destroy_process_group()
