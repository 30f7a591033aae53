"""Joins the per-instruction counters of an ncu report (SASS source page) with nvdisasm's line info of the same kernel,
and prints the instruction / stall-sample share per source line (development aid).
Usage: python scratch/ncu_lines.py report.ncu-rep mangled_kernel_substring [top_n [cubin_prefix source_file]]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "lle_b200", "_native", "liblle_b200.so")], cwd=tmp, capture_output=True)
cubin_prefix = sys.argv[4] if len(sys.argv) > 4 else "vec_world"
source_file = sys.argv[5] if len(sys.argv) > 5 else "world_kernel.cuh"
cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_prefix)][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines_of = []  # program order: (offset, line, text)
inside, cur = False, None
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?", line (\d+)', ln)
    if m:
        cur = int(m.group(1))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines_of.append((int(m.group(1), 16), cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ie, sm = h.index("Instructions Executed"), h.index("# Samples")
data = rows[2:]
assert len(data) == len(lines_of), (len(data), len(lines_of))
src = open(os.path.join(root, "lle_b200", "csrc", source_file)).read().splitlines()
inst, samp = collections.Counter(), collections.Counter()
for r, (off, line, text) in zip(data, lines_of):
    inst[line] += float(r[ie] or 0)
    samp[line] += float(r[sm] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
# coarse regions of world_kernel.cuh (by the line a SASS instruction is attributed to; inlined helpers count where they are defined)
# per-function view: every `__device__` helper / method and the kernel's commented sections start a region
import re as _re
fmarks = [(1, "file head")]
for n, text in enumerate(src, 1):
    m = _re.search(r"(?:__device__ __forceinline__|LLE_HD(?:_NOINLINE)?) [\w:<>\*&\s]+?\b(\w+)\(", text)
    if m:
        fmarks.append((n, m.group(1)))
    elif "__global__" in text and "lle_world_kernel" in text:
        fmarks.append((n, "kernel: prologue"))
    elif "// ====" in text and "logic" in text:
        fmarks.append((n, "kernel: logic glue + small outputs"))
    elif "// ====" in text and "observations of the group" in text:
        fmarks.append((n, "kernel: render"))
fmarks.append((len(src) + 1, "end"))
rows_f = []
for (a, name), (b, _) in zip(fmarks, fmarks[1:]):
    vi = sum(v for l, v in inst.items() if l and a <= l < b)
    vs = sum(v for l, v in samp.items() if l and a <= l < b)
    if vi or vs:
        rows_f.append((vi / ti * 100, vs / ts * 100, name, a))
print("function / section      instr%  stall%")
for vi, vs, name, a in sorted(rows_f, reverse=True)[:22]:
    print(f"  {name:34s} {vi:5.1f}  {vs:5.1f}   (line {a})")
marks = []
for n, text in enumerate(src, 1):
    for tag in ("struct World {", "struct PatchCache", "template <int MODE, bool FAST>", "// ================================================================== logic",
                "// ================================================================== observations of the group", "} else {\n"):
        if tag in text:
            marks.append((n, text.strip()[:60]))
marks.append((len(src) + 1, "end"))
prev = (1, "helpers / PTX wrappers")
for m in marks:
    a, b = prev[0], m[0]
    vi = sum(v for l, v in inst.items() if l and a <= l < b)
    vs = sum(v for l, v in samp.items() if l and a <= l < b)
    print(f"region {a:5d}-{b - 1:5d}  instr {vi / ti * 100:5.1f}%  stall {vs / ts * 100:5.1f}%   {prev[1]}")
    prev = m
print("line   instr%  stall%  source")
for line, v in sorted(inst.items(), key=lambda x: -(x[1] / ti + samp[x[0]] / ts))[:top_n]:
    print(f"{line:5d}  {v / ti * 100:5.1f}  {samp[line] / ts * 100:5.1f}   {src[line - 1].strip()[:120] if line else ''}")
