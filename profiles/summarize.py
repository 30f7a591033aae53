"""Regenerates the tracked ncu summaries in profiles/ from gpurun_out/ (scratch, untracked).
Usage: python profiles/summarize.py [round_tag]"""
import collections, csv, json, os, statistics, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
CMD = "python bench.py --steps 64 --warmup 8 --no-cpu-baseline --e2e-steps 8 --no-configs --no-compiled-host"
if tag == "r02f":  # the final capture of round 2 (profiles/capture_final.sh): the closed loop over parts cannot run under ncu
    CMD += " --no-closed-loop"

rows = list(csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))))
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
by = collections.defaultdict(list)
for d in data:
    if d["Metric Name"] == "gpu__time_duration.sum":
        by[d["Kernel Name"]].append(float(d["Metric Value"]))
tot = sum(sum(v) for v in by.values())
lines = [f"# ncu launch list, {tag} - `{CMD}`",
         "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 (cold-cache, serialised: compare SHARES, not absolutes)",
         "kernel,launches,total_us,mean_us,share"]
for k, v in sorted(by.items(), key=lambda x: -sum(x[1])):
    lines.append(f"\"{k[:110]}\",{len(v)},{sum(v) / 1e3:.1f},{statistics.mean(v) / 1e3:.2f},{sum(v) / tot:.4f}")
open(os.path.join(ROOT, "profiles", f"launches_{tag}_summary.csv"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:6]))

out = subprocess.run(["ncu", "-i", os.path.join(G, f"prof_{tag}_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sm__cycles_elapsed.avg"]
summary = {}
txt = [f"# ncu --set full --clock-control none --import-source on -k regex:lle_world_kernel -s 40 -c 3, {tag}",
       f"# command: {CMD}   (65,536 x level 6; step kernel lle_world_kernel<MODE_STEP, FAST>; under ncu launches are serialised,",
       "# so the cross-launch overlap of the real run (programmatic dependent launch + epoch flags) is absent here)",
       "metric,unit,launch1,launch2,launch3"]
for name in want:
    if name in hdr:
        i = hdr.index(name)
        vals = [r[i] for r in rows[2:]]
        txt.append(f"{name},{units[i]}," + ",".join(vals))
        summary[name] = (units[i], vals)
open(os.path.join(ROOT, "profiles", f"ncu_full_{tag}_summary.csv"), "w").write("\n".join(txt) + "\n")
print("\n".join(txt[3:9]))
mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
ur, rd = summary["dram__bytes_read.sum"]
uw, wr = summary["dram__bytes_write.sum"]
traffic = statistics.mean(float(r) * mult[ur] + float(w) * mult[uw] for r, w in zip(rd, wr))
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
json.dump({"dram_bytes_per_launch": traffic, "commit": commit,
           "source": f"profiles/ncu_full_{tag}_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum, mean of 3 launches)",
           "note": "below the 501 MB algorithmic figure because the tail of a launch's writes is still in the 126 MB L2 when the kernel "
                   "ends (written back during the next launch); no re-reads: dram read is 2 MB per launch"},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("traffic", traffic)
