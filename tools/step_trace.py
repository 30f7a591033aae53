"""Per-step device time of the first steps after an idle period (is the 20-step window of the driver's bench run slower
than steady state because of clock ramp-up, launch-queue depth or the grid choice?).

    python tools/step_trace.py [--config 2] [--steps 200] [--idle-ms 500]

Prints one JSON line: ms of step k (CUDA events between consecutive launches) for k < steps, after (a) creation + idle,
(b) a 300 ms preheat + device sync.  Events between launches serialise nothing (they are recorded on the same stream),
but they do sit between the programmatically dependent launches, so the absolute values are upper bounds."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lle_b200 import workloads

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--idle-ms", type=float, default=500)
args = ap.parse_args()
wl = workloads.build(args.config)


def window(k):
    """ms of k back-to-back steps (one pair of events around the window)."""
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        wl.step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


out = {"config": args.config}
time.sleep(args.idle_ms / 1e3)
out["cold_5+20"] = [window(5), window(20), window(20), window(20)]
time.sleep(args.idle_ms / 1e3)
out["cold_windows_of_20"] = [window(20) for _ in range(10)]
t0 = time.perf_counter()
while time.perf_counter() - t0 < 0.3:
    for _ in range(64):
        wl.step()
    torch.cuda.synchronize()
out["hot_windows_of_20"] = [window(20) for _ in range(10)]
out["hot_windows_of_200"] = [window(200) for _ in range(5)]
out["hot_windows_of_2000"] = [window(2000) for _ in range(3)]
print(json.dumps(out))
