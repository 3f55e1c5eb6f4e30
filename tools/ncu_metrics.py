#!/usr/bin/env python
"""Print the headline metrics of every launch in an ncu report (the text that goes under profiles/).

    python tools/ncu_metrics.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
for r in rows[2:]:
    print("Kernel Name =", r[h.index("Kernel Name")])
    for w in WANT:
        if w in h:
            i = h.index(w)
            print(f"{w} = {r[i]} {units[i]}")
    stalls = [(h[i].split("smsp__average_warps_issue_stalled_")[1].split("_per_issue")[0], float(r[i] or 0)) for i in range(len(h))
              if h[i].startswith("smsp__average_warps_issue_stalled_") and h[i].endswith("_per_issue_active.ratio")]
    if stalls:
        tot = sum(v for _, v in stalls)
        print("warp stall mix (per issue active): " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(stalls, key=lambda x: -x[1])[:8]))
    print("---")
