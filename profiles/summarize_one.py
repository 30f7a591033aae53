"""Tracked summary of one ncu capture: python profiles/summarize_one.py gpurun_out/<report>.ncu-rep profiles/<out>.csv "<command / note>" """
import csv, subprocess, sys

rep, out_path, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
txt = [f"# ncu --set full --clock-control none --import-source on, one launch; {note}", "metric,unit,value"]
for name in want:
    if name in hdr:
        i = hdr.index(name)
        txt.append(f"{name},{units[i]},{rows[2][i]}")
open(out_path, "w").write("\n".join(txt) + "\n")
print("\n".join(txt[:12]))
