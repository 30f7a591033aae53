"""Per-kernel SASS evidence of the built library: registers / shared memory / spills (cuobjdump -res-usage) and counts of the
mnemonics that matter here — UBLKCP (TMA 1-D bulk copy), LDGSTS (cp.async), SYNCS (mbarrier), LDL / STL (local memory),
MEMBAR, ATOMG / RED, UTMA* / UTC* (tensor-map TMA / tcgen05: expected 0, nothing on this path is a contraction).
    python tools/sass_summary.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lle_b200", "_native", "liblle_b200.so")
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    if name and "REG:" in line:
        usage[name] = line.strip()
        name = None
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
arch = set()
cur = None
WATCH = ("UBLKCP", "LDGSTS", "SYNCS", "LDL", "STL", "MEMBAR", "ATOMG", "RED", "UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "HMMA", "STG", "LDG", "STS", "LDS", "NANOSLEEP", "SHFL", "VOTE")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        if op in WATCH:
            counts[cur][op] += 1


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:110]


print(f"# SASS summary of lle_b200/_native/liblle_b200.so at commit {head}; architectures in the fat binary: {sorted(arch)}")
print("# cuobjdump -sass / -res-usage; counts are static instruction counts per kernel")
for k in sorted(counts, key=lambda k: -counts[k]["_total"]):
    c = counts[k]
    print(f"\n{demangle(k)}")
    print(f"  {usage.get(k, '')}")
    print("  instructions " + str(c["_total"]) + "  " + "  ".join(f"{op}={c[op]}" for op in WATCH if c[op]))
