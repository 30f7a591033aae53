"""Times the layout-generator kernel (lle_gen_run) with CUDA events: attempts/s for a few configurations, next to the same
generator core compiled for the host (one thread) and the Python oracle.  Development aid; prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from lle_b200.generator import WorldGenerator  # noqa: E402

CASES = [
    ("5x5_a2_l2 (config 3)", dict(width=5, height=5, n_agents=2, n_lasers=2), 1 << 20),
    ("10x10_a4_l4_g8_shapes", dict(width=10, height=10, n_agents=4, n_lasers=4, n_gems=8, walls_style="shapes"), 1 << 19),
    ("lanes 8x8 cross-agent", dict(width=8, height=8, n_agents=3, starts="edge", exits="opposite", n_lasers=2, laser_placement="cross-agent"), 1 << 19),
    ("32x32_a4_l4", dict(width=32, height=32, n_agents=4, n_lasers=4, n_gems=8), 1 << 17),
]


def main():
    from oracle import generator as og
    from test_generator import HostCore

    host = HostCore()
    for name, kw, n in CASES:
        g = WorldGenerator(**kw, batch=n)
        for _ in range(2):
            g.run(first_seed=0, n=n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for r in range(reps):
            cells, status, labels, tries = g.run(first_seed=r * n, n=n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        accepted = float(status.float().mean())
        if os.environ.get("GEN_BENCH_GPU_ONLY"):
            print(json.dumps({"case": name, "attempts": n, "ms": round(ms, 3), "attempts_per_s": round(n / ms * 1e3), "accepted": round(accepted, 4)}))
            del g
            continue
        cfg = og.GenConfig(**kw)
        m = 20000 if kw["width"] <= 10 else 2000
        t = time.perf_counter()
        host.run(cfg, list(range(m)))
        host_rate = m / (time.perf_counter() - t)
        m2 = 300
        t = time.perf_counter()
        for s in range(m2):
            og.try_generate(cfg, s)
        py_rate = m2 / (time.perf_counter() - t)
        print(json.dumps({"case": name, "attempts": n, "ms": round(ms, 3), "attempts_per_s": round(n / ms * 1e3), "accepted": round(accepted, 4),
                          "host_core_1thread_attempts_per_s": round(host_rate), "python_oracle_attempts_per_s": round(py_rate)}))
        del g


if __name__ == "__main__":
    main()
