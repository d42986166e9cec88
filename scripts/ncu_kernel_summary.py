"""Text summary of an .ncu-rep (run here, no GPU): per captured launch the headline metrics, the stall-reason breakdown
and the hottest SASS instructions by stall samples. Usage: ncu_kernel_summary.py <file.ncu-rep> [launch index ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
pick = [int(a) for a in sys.argv[2:]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for li, r in enumerate(rows[2:]):
    if pick and li not in pick:
        continue
    d = dict(zip(hdr, r))
    print(f"=== launch {li}: {d.get('Kernel Name')}")
    for k in KEYS:
        if k in d:
            print(f"{k:84s} {d[k]:>18s} {units[hdr.index(k)]}")
    st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(d[h]))
          for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and d[h]]
    tot = sum(v for _, v in st) or 1.0
    print("warp stall reasons (share of stalled + issuing warp slots per issue):")
    for n, v in sorted(st, key=lambda x: -x[1])[:9]:
        print(f"    {n:24s} {v:7.3f}  ({100 * v / tot:4.1f}%)")
    if "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed" in d and "sm__cycles_elapsed.max" in d:
        cyc = float(d["sm__cycles_elapsed.max"])
        f64 = sum(float(d[f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed"]) for o in ("dfma", "dmul", "dadd")) * cyc
        print(f"FP64 thread instructions (dfma + dmul + dadd, predicated on): {f64:.4e}")
