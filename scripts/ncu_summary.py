#!/usr/bin/env python
"""ncu report -> text summary (key raw metrics + hottest source lines): python scripts/ncu_summary.py rep.ncu-rep out.txt"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__maximum_warps_per_active_cycle_pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "launch__shared_mem_per_block_dynamic", "derived__local_spilling_requests"]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none, report {rep.split('/')[-1]}\n")
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        for k in want:
            if k in h:
                i = h.index(k)
                f.write(f"{k} [{u[i]}] = {r[i]}\n")
        f.write("\n")
    lines = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, "30"], capture_output=True, text=True).stdout
    f.write("# hottest source lines (stall samples, share of warp instructions, active threads per instruction)\n" + lines)
print(open(out).read()[:1500])
