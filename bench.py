#!/usr/bin/env python
"""bench.py — LF-BA hot-path benchmark (contract: one JSON line on rank 0).

Metric (BASELINE.json): "LM iters/s & M residual+Jacobian evals/s, 1/2/4/8 B200 vs Ceres CPU".
  metric  = M residual+Jacobian evals/s over complete LM solves (every LM iteration makes exactly one fused
            residual+Jacobian pass over all observations), `lm_iters_per_s` is reported next to it.
  step    = one complete LM solve (lfba_solver_run) of the workload from the same initial guess.
  value   = device-resident: observations already indexed in HBM, timed with CUDA events inside the library on its
            own stream (summary.solve_gpu_ms), max over ranks.
  e2e     = the same metric through the drop-in call lfba_solve() with HOST buffers: H2D of all observation arrays
            (pinned), device indexing, LM solve, D2H of camera/views/points — everything inside the timed region.
Workload: BASELINE.json configs[3] — 1M points x 1000 frames (window 4), ~1.1e8 micro-image observations, 2 radial +
tangential, robust (Cauchy 0.5), poses + points refined, micro-lens-centre adjustment. It fits one B200 (~10 GB), so it
is the N=1 workload too; with N GPUs the same scene is sharded by point/frame range (strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg3|cfg2|cfg1]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the ncu --set full
# capture of the same build (profiles/README.md); None where no capture exists for that workload
NCU_TRAFFIC = {"cfg4": 4.397e9}  # 3.262 GB read + 1.135 GB written per launch (profiles/r01_ncu_k_eval_rows_cfg4.txt)
FP64_INST_PER_OBS = 262.4  # ncu source counters of k_eval_rows<9,2,1> (profiles/README.md): loop + per-track part / N
METRIC = "lm_residual_jacobian_evals_per_s"
UNIT = "M evals/s"
WORKLOADS = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4}
WORKLOAD_DESC = {
    "cfg1": "calib_marker 500 pts x 10 frames, 3 distance constraints",
    "cfg2": "recalib 5k pts x 20 frames (fL,B fixed, bounds)",
    "cfg3": "full calibration 50k pts x 100 frames, window 20",
    "cfg4": "scaled scene 1M pts x 1000 frames, window 4, ~1.1e8 observations",
}


# ---------------------------------------------------------------------------------------------------
# helpers shared with tests/test_sharding_gloo.py
# ---------------------------------------------------------------------------------------------------
def shard_range(n_points: int, rank: int, world: int):
    """Contiguous point range of a rank. Points are generated sorted by the start of their visibility window, so a
    point range is a frame range; every observation of a point stays on the point's rank."""
    lo = (n_points * rank) // world
    hi = (n_points * (rank + 1)) // world
    return lo, hi


def broadcast_unique_id(uid: bytes | None, device="cuda") -> bytes:
    """Rank 0 creates the 128-byte ncclUniqueId (lfba_comm_unique_id); torch.distributed is only the plumbing."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t.copy_(torch.tensor(list(uid), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                pw.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def pinned_problem(pa):
    """Copies of the problem arrays in pinned host memory (torch is plumbing: it only owns the page-locked pages)."""
    import torch
    from lifcal_b200 import capi
    keep = []

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keep.append(t)
        return t.numpy()
    out = capi.ProblemArrays.__new__(capi.ProblemArrays)
    out.__dict__.update(pa.__dict__)
    for k in ("obs_x", "obs_y", "ml_x", "ml_y", "point_idx", "frame_idx"):
        setattr(out, k, pin(getattr(pa, k)))
    out._keep = keep
    return out


# ---------------------------------------------------------------------------------------------------
# CPU baseline (oracle) on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------
def cpu_sample_spec(workload: str, small: bool):
    """A bounded sample of the workload with the same structure (window, flags, observations per view)."""
    from lifcal_b200 import capi
    preset = WORKLOADS[workload]
    if workload == "cfg4":
        n_points, n_frames = (6000, 12) if small else (20000, 24)
        return capi.scene_spec(preset, n_points=n_points, n_frames=n_frames, order=1), \
            f"{n_points} points x {n_frames} frames of cfg4 (window 4, same flags)"
    if workload == "cfg3":
        n_points, n_frames = (2000, 40) if small else (5000, 60)
        return capi.scene_spec(preset, n_points=n_points, n_frames=n_frames), \
            f"{n_points} points x {n_frames} frames of cfg3 (window 20, same flags)"
    if workload == "cfg2":
        n_points = 1000 if small else 5000
        return capi.scene_spec(preset, n_points=n_points), f"{n_points} points x 20 frames of cfg2"
    return capi.scene_spec(preset), "the full cfg1 scene"


def run_cpu_baseline(workload: str, small: bool, steps: int = 1, warmup: int = 0):
    from lifcal_b200 import capi
    from oracle import binding as ob
    spec, desc = cpu_sample_spec(workload, small)
    sc = capi.Scene(spec)
    # all host cores this process may use — torchrun exports OMP_NUM_THREADS=1, which would silently make the reference
    # arm single-threaded at N > 1
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    threads = max(ob.max_threads(), ncores)
    times, evals, iters = [], 0, 0
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, _, s = ob.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init, threads=threads)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
            evals += s["num_jacobian_evals"]
            iters += s["num_iterations"]
    total = sum(times)
    n = sc.problem.n_obs
    return {"value": n * evals / total / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{desc}: {n} observations, {iters // max(1, steps)} LM iterations/solve, "
                      f"{total / max(1, steps):.2f} s/solve; oracle = Ceres-2.1.0-equivalent restatement "
                      f"(Jet<26> autodiff, DENSE_SCHUR, dense LLT), functor pinned bit-exact to the reference headers",
            "lm_iters_per_s": iters / total, "seconds": total, "n_obs": n, "steps": steps}


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("LFBA_BENCH_WORKLOAD", "cfg4"), choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: {WORKLOAD_DESC[args.workload]}", "flags": "nRadial=2,tangential,robust(Cauchy 0.5),"
              "refinePoses,refinePoints,mlAdj", "l2": "inputs larger than L2 (no flush needed)"}

    # ------------------------------------------------------------------ reference arm: CPU oracle
    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = run_cpu_baseline(args.workload, small=True, steps=max(1, args.steps), warmup=min(1, args.warmup))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": min(1, args.warmup), "ms_per_step": 1e3 * cb["seconds"] / max(1, args.steps),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config, sample=cb["sample"]), "lm_iters_per_s": cb["lm_iters_per_s"],
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    from lifcal_b200 import api, capi
    if api.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the LF-BA path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    # ---- synthetic scene: every rank generates its own shard (counter-based RNG) ----
    preset = WORKLOADS[args.workload]
    spec = capi.scene_spec(preset, order=1)
    P = spec.n_points
    if world > 1 and spec.n_constraints > 0:
        spec.n_constraints = 0  # constraint-coupled points would all go to rank 0; the scaling workload has none
    lo, hi = shard_range(P, rank, world)
    if world > 1:
        spec.point_begin, spec.point_end = lo, hi
    t_gen = time.perf_counter()
    sc = capi.Scene(spec)
    t_gen = time.perf_counter() - t_gen
    pa = sc.problem
    n_local = pa.n_obs
    opt = api.default_options(device=local_rank)

    comm = None
    if world > 1:
        uid = broadcast_unique_id(api.comm_unique_id() if rank == 0 else None)
        comm = api.Communicator(rank, world, uid)  # one NCCL communicator per rank for the whole run
    t_setup = time.perf_counter()
    ds = api.DeviceSolver(pa, opt, communicator=comm)
    ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    s = None
    for _ in range(max(3, args.warmup)):
        s = ds.run()
    n_global = s["num_observations"]
    # ---- timed: exactly K solves ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    gpu_ms, evals, iters, launches = 0.0, 0, 0, 0
    for _ in range(args.steps):
        s = ds.run()
        gpu_ms += s["solve_gpu_ms"]
        evals += s["num_jacobian_evals"]
        iters += s["num_iterations"]
        launches += s["gpu_launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([gpu_ms, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gpu_ms, wall = float(t[0].item()), float(t[1].item())
    value = n_global * evals / (gpu_ms * 1e-3) / 1e6

    # ---- dominant kernel: fused eval, timed live with CUDA events on the library's stream ----
    # Algorithmic HBM bytes of one launch: 20 B per observation (double2 + int32 of the packed
    # stream) + the per-track record written, (9 + 3 NC) * 8 B = 288 B per (point, frame) track.
    eval_ms = ds.time_eval(reps=10, materialize=False)
    rec_stride = 9 + 3 * 9
    alg_bytes = 20.0 * n_local + 8.0 * rec_stride * s["num_tracks"]
    peak, peak_src = measured_peaks()
    roofline = {"bound": "hbm", "kernel": "k_eval_rows (fused residual + analytic Jacobian + Cauchy weighting + per-track Gram "
                "-> normal-equation blocks; Jacobian never leaves registers)",
                "achieved": alg_bytes / (eval_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg_bytes / (eval_ms * 1e-3) / 1e9 / peak, "traffic": NCU_TRAFFIC.get(args.workload),
                "peak_source": peak_src, "ms_per_launch": eval_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "note": "this kernel is FP64-pipe bound by design (30 B/observation of HBM traffic against 263 FP64 "
                        "instructions): its fraction of the HBM roof is small on purpose; see fp64 for the pipe "
                        "utilisation and roofline_eval_only for the HBM-bound kernel that materialises the Jacobian"}
    extra = {}
    if rank == 0 and world == 1:
        try:
            mat_ms = ds.time_eval(reps=3, materialize=True)
            mat_bytes = (28.0 + 16.0 + 16.0 * 26.0) * n_local
            extra["roofline_eval_only"] = {"bound": "hbm", "kernel": "k_eval_only (residual + Jacobian materialised in Ceres' "
                                           "block layout, 2x(17+6+3) doubles per observation)",
                                           "achieved": mat_bytes / (mat_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                           "frac": mat_bytes / (mat_ms * 1e-3) / 1e9 / peak, "ms_per_launch": mat_ms,
                                           "algorithmic_bytes_per_launch": mat_bytes,
                                           "traffic": 5.1557e10 if args.workload == "cfg4" else None,  # ncu: 3.17 GB read + 48.39 GB written
                                           "m_evals_per_s": n_local / (mat_ms * 1e-3) / 1e6}
            fp64 = api.measure_fp64_peak(local_rank)
            # FP64 instructions per observation of the fused kernel, from the ncu source counters of the same build
            # (profiles/README.md): per-observation loop + the per-track expansion amortised over the track
            fp64_inst_per_obs = FP64_INST_PER_OBS
            ach = fp64_inst_per_obs * n_local / (eval_ms * 1e-3) * 2.0 / 1e12  # counted as 2 flop per FP64 instruction
            extra["fp64"] = {"bound": "fp64", "measured_dfma_peak_tflops": fp64, "fp64_inst_per_observation": fp64_inst_per_obs,
                             "achieved_tflops_equiv": ach, "frac": ach / fp64,
                             "fused_eval_m_evals_per_s": n_local / (eval_ms * 1e-3) / 1e6}
        except Exception as e:  # noqa: BLE001
            extra["roofline_eval_only"] = {"error": str(e)}
    ds.close()

    # ---- e2e: the drop-in call with host (pinned) buffers, H2D + indexing + solve + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        ppa = pinned_problem(pa)
        o2 = api.default_options(device=local_rank)
        ts, ev2 = [], 0
        if world == 1:
            for k in range(1 + min(2, args.steps)):
                barrier()
                t1 = time.perf_counter()
                cam, vw, pt, s2 = api.solve(ppa, sc.camera_init, sc.views_init, sc.points_init, o2)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t1
                if k > 0:
                    ts.append(dt)
                    ev2 += s2["num_jacobian_evals"]
        else:
            for k in range(1 + min(2, args.steps)):
                barrier()
                t1 = time.perf_counter()
                d2 = api.DeviceSolver(ppa, o2, communicator=comm)
                d2.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
                s2 = d2.run()
                d2.get_parameters()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t1
                d2.close()
                tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                if k > 0:
                    ts.append(float(tt.item()))
                    ev2 += s2["num_jacobian_evals"]
        tot = sum(ts)
        F = pa.n_frames
        e2e = {"value": n_global * ev2 / tot / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(40 * n_local + 8 * (17 + 6 * F + 3 * P)),
               "d2h_bytes_per_step": int(8 * (17 + 6 * F + 3 * P)), "s_per_solve": tot / len(ts),
               "api": "lfba_solve (C ABI, pinned host buffers)" if world == 1 else "lfba_solver_create+run per rank"}

    if comm is not None:
        comm.close()
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": gpu_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "lm_iters_per_s": iters / (gpu_ms * 1e-3),
            "lm_iterations_per_solve": iters / args.steps, "observations": n_global, "tracks_rank0": s["num_tracks"],
            "lenses_rank0": s["num_lenses"], "reduced_system_size": s["reduced_system_size"],
            "wall_s_per_step": wall / args.steps, "final_cost": s["final_cost"], "scene_gen_s": t_gen, "setup_s": t_setup,
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "e2e": e2e}
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = {k: v for k, v in run_cpu_baseline(args.workload, small=False).items()
                                if k in ("value", "unit", "cores", "kind", "sample", "lm_iters_per_s")}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
