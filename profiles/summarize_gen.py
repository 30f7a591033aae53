"""Tracked summary of the layout-generator kernel's ncu capture (gpurun_out/prof_gen_<tag>.ncu-rep, scratch) ->
profiles/gen_ncu_<tag>_summary.csv.  Usage: python profiles/summarize_gen.py [tag]"""
import csv, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", f"prof_gen_{tag}.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
txt = [f"# ncu --set full --clock-control none --import-source on -k regex:lle_gen_kernel -s 1 -c 1, {tag}",
       "# command: python scratch/gen_prof.py   (5x5, 2 agents, 2 lasers, 148 * 92 * 8 = 108,928 attempts in the profiled launch)",
       "metric,unit,value"]
for name in want:
    if name in hdr:
        i = hdr.index(name)
        txt.append(f"{name},{units[i]},{rows[2][i]}")
open(os.path.join(ROOT, "profiles", f"gen_ncu_{tag}_summary.csv"), "w").write("\n".join(txt) + "\n")
print("\n".join(txt))
