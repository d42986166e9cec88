"""CPU test of the multi-process plumbing used by bench.py under torchrun (world_size 2, gloo): the NCCL unique id
is broadcast as bytes, every rank generates its own shard of the scene, shards are disjoint and complete, and
every observation of a point lives on exactly one rank."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import bench
    from lifcal_b200 import capi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = bench.broadcast_unique_id(bytes(range(128)) if rank == 0 else None, device="cpu")
    lo, hi = bench.shard_range(1000, rank, world)
    sc = capi.make_scene(None, n_points=1000, n_frames=30, window=4, seed=9, order=1, point_begin=lo, point_end=hi)
    n = torch.tensor([sc.problem.n_obs], dtype=torch.int64)
    dist.all_reduce(n)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = dict(uid=uid, lo=lo, hi=hi, n=sc.problem.n_obs, total=int(n.item()), tmax=float(t.item()),
                     pmin=int(sc.problem.point_idx.min()), pmax=int(sc.problem.point_idx.max()))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo(built):
    from lifcal_b200 import capi
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["uid"] == r1["uid"] == bytes(range(128))
    assert r0["lo"] == 0 and r0["hi"] == r1["lo"] and r1["hi"] == 1000
    full = capi.make_scene(None, n_points=1000, n_frames=30, window=4, seed=9, order=1)
    assert r0["total"] == r1["total"] == full.problem.n_obs == r0["n"] + r1["n"]
    assert r0["pmax"] < r1["pmin"]  # a point's observations never straddle ranks
    assert r0["tmax"] == 2.0
