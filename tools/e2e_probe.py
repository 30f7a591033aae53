"""Where does the closed-loop end-to-end time go?  Host-side timings (perf_counter_ns) of every call of the loop bench.py
runs in its e2e arm — submit, wait, the policy's read of the results — for one whole-batch env and for K sub-batches in
flight, plus the isolated kernel time of a sub-batch.

    python tools/e2e_probe.py [--steps 300] [--parts 1 2 4]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import lle_b200

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--parts", type=int, nargs="+", default=[1, 2, 4])
args = ap.parse_args()
N, K = args.envs, args.steps
T0 = 10_000_000
full = lle_b200.VecWorld(lle_b200.Map(level=6), N, seed=1)
A, R = full.n_agents, full.reward_dim
full.step_count = T0
rec_act = torch.empty((K, N, A), dtype=torch.int8).pin_memory()
rec_done = torch.empty((K, N), dtype=torch.uint8).pin_memory()
for s in range(K):
    full.step(None)
    rec_act[s].copy_(full.actions, non_blocking=True)
    rec_done[s].copy_(full.done, non_blocking=True)
torch.cuda.synchronize()
del full
out = {}
for parts in args.parts:
    n = N // parts
    vecs = [lle_b200.VecWorld(lle_b200.Map(level=6), n, seed=1, env_id_base=h * n) for h in range(parts)]
    acts = [rec_act[:, h * n:(h + 1) * n].contiguous().pin_memory() for h in range(parts)]
    dones = [rec_done[:, h * n:(h + 1) * n].contiguous().pin_memory() for h in range(parts)]
    rw = [torch.empty((n, R), dtype=torch.float32).pin_memory() for _ in range(parts)]
    dn = [torch.empty((n,), dtype=torch.uint8).pin_memory() for _ in range(parts)]
    dn_np, dones_np = [t.numpy() for t in dn], [t.numpy() for t in dones]
    for rep in range(2):
        for v in vecs:
            v.reset()
            v.step_count = T0
        torch.cuda.synchronize()
        t_submit = t_wait = t_policy = 0
        bad = 0
        t0 = time.perf_counter_ns()
        for s in range(K):
            for h in range(parts):
                a = acts[h][s]
                if s > 0:
                    c0 = time.perf_counter_ns()
                    vecs[h].wait_host()
                    c1 = time.perf_counter_ns()
                    if not np.array_equal(dn_np[h], dones_np[h][s - 1]):
                        bad += 1
                    c2 = time.perf_counter_ns()
                    t_wait += c1 - c0
                    t_policy += c2 - c1
                c0 = time.perf_counter_ns()
                vecs[h].submit_host(a, rw[h], dn[h], after_current_stream=False)
                t_submit += time.perf_counter_ns() - c0
        for h in range(parts):
            vecs[h].wait_host()
        torch.cuda.synchronize()
        total = time.perf_counter_ns() - t0
    calls = K * parts
    # isolated kernel time of one sub-batch: device-side step, synchronised every launch
    v = vecs[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iso = []
    for _ in range(50):
        e0.record()
        v.step(None)
        e1.record()
        torch.cuda.synchronize()
        iso.append(e0.elapsed_time(e1) * 1e3)
    out[str(parts)] = {"envs_per_part": n, "us_per_full_step": total / K / 1e3, "env_steps_per_s": N * K / (total / 1e9), "mismatches": bad,
                       "host_us_per_call": {"submit": t_submit / calls / 1e3, "wait": t_wait / max(calls - parts, 1) / 1e3,
                                            "policy_read": t_policy / max(calls - parts, 1) / 1e3},
                       "isolated_kernel_us": sorted(iso)[len(iso) // 2]}
    del vecs
print(json.dumps(out))
