"""Joins the per-instruction counters of an ncu report (SASS source page) with nvdisasm's line info of the same kernel, and
prints the instruction / stall-sample share per source FILE, per function and per source line (development aid).

    python tools/ncu_lines.py report.ncu-rep mangled_kernel_substring [top_n [cubin_prefix]]

The report must come from the library as it is built now (same SASS): instruction k of the report is instruction k of the
disassembly.  Inlined helpers count where they are defined."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cubin_prefix = sys.argv[4] if len(sys.argv) > 4 else "vec_world"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "lle_b200", "_native", "liblle_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_prefix)][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
where = []  # program order: (file, line)
inside, cur = False, (None, 0)
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln):
        where.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ie, sm = h.index("Instructions Executed"), h.index("# Samples")
data = [r for r in rows[2:] if len(r) == len(h) and r[0].strip().startswith("0x")]
if len(data) > len(where) and len(where):  # several launches in the report (each with its own header rows): keep the first one
    data = data[:len(where)]
assert len(data) == len(where), f"the report has {len(data)} instructions, the current build {len(where)}: rebuild the library the report was taken with"
inst, samp = collections.Counter(), collections.Counter()
for r, key in zip(data, where):
    inst[key] += float(r[ie] or 0)
    samp[key] += float(r[sm] or 0)
ti, ts = sum(inst.values()), max(sum(samp.values()), 1.0)
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
files = sorted({f for f, _ in inst if f})
sources = {}
for f in files:
    path = os.path.join(root, "lle_b200", "csrc", f)
    sources[f] = open(path).read().splitlines() if os.path.exists(path) else []
print("file                      instr%  stall%")
for f in files:
    print(f"  {f:24s} {sum(v for (ff, _), v in inst.items() if ff == f) / ti * 100:5.1f}  {sum(v for (ff, _), v in samp.items() if ff == f) / ts * 100:5.1f}")
print("function / section                                  instr%  stall%")
rows_f = []
for f in files:
    src = sources[f]
    marks = [(1, "file head")]
    for n, text in enumerate(src, 1):
        m = re.search(r"(?:__device__ __forceinline__|LLE_HD(?:_NOINLINE)?) [\w:<>\*&\s]+?\b(\w+)\(", text)
        if m:
            marks.append((n, m.group(1)))
        elif "__global__" in text:
            marks.append((n, "kernel prologue"))
        elif text.strip().startswith("// ----") or text.strip().startswith("// ===="):
            marks.append((n, text.strip().strip("/=- ")[:40]))
    marks.append((len(src) + 1, "end"))
    for (a, name), (b, _) in zip(marks, marks[1:]):
        vi = sum(v for (ff, l), v in inst.items() if ff == f and a <= l < b)
        vs = sum(v for (ff, l), v in samp.items() if ff == f and a <= l < b)
        if vi or vs:
            rows_f.append((vi / ti * 100, vs / ts * 100, f"{f}:{a} {name}"))
for vi, vs, name in sorted(rows_f, reverse=True)[:30]:
    print(f"  {name:50s} {vi:5.1f}  {vs:5.1f}")
print("line                           instr%  stall%  source")
for key, v in sorted(inst.items(), key=lambda x: -(x[1] / ti + samp[x[0]] / ts))[:top_n]:
    f, line = key
    src = sources.get(f, [])
    text = src[line - 1].strip()[:110] if 0 < line <= len(src) else ""
    print(f"{(f or '?') + ':' + str(line):30s} {v / ti * 100:5.1f}  {samp[key] / ts * 100:5.1f}   {text}")
