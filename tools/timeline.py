"""Per-warp timeline of ONE step launch (development aid, LLE_B200_TIMELINE=1): when the warps start, issue their first and last
observation store, and end — for a launch that overlaps its predecessors (free-running) or an isolated one (device idle before).
    python tools/timeline.py [--level 6] [--envs 65536] [--isolated]"""
import argparse
import ctypes as C
import json
import os
import sys

os.environ["LLE_B200_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import lle_b200
from lle_b200._native import lib

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=6)
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--isolated", action="store_true")
args = ap.parse_args()
vec = lle_b200.VecWorld(lle_b200.Map(level=args.level), args.envs, seed=1)
for _ in range(20):
    vec.step(None)
if args.isolated:
    torch.cuda.synchronize()
vec.step(None)
n = C.c_int64(0)
lib().lle_vec_debug_timeline(vec._h, None, 0, C.byref(n))
buf = np.zeros((n.value, 4), dtype=np.uint64)
lib().lle_vec_debug_timeline(vec._h, buf.ctypes.data, n.value, C.byref(n))
t = buf.astype(np.int64)
act = t[:, 1] > 0
t0 = t[t[:, 0] > 0, 0].min()
rel = (t - t0) / 1e3
q = lambda x: [round(float(v), 1) for v in np.percentile(x, [0, 10, 50, 90, 100])]
print(json.dumps({"envs": args.envs, "isolated": args.isolated, "warps": int(n.value), "warps_with_work": int(act.sum()),
                  "start_us_p0_10_50_90_100": q(rel[act, 0]), "first_store_us": q(rel[act, 1]), "last_store_us": q(rel[act, 2]),
                  "end_us": q(rel[act, 3]), "store_span_us_per_warp": q(rel[act, 2] - rel[act, 1])}))
