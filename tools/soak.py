"""Soak of the dataflow-ordered stepping: the same batch advanced by (a) single-step launches overlapped with programmatic
dependent launch, (b) rollout launches of 128 steps, (c) rollout launches of 7 steps interleaved with single steps, for many
steps; the three must end in bit-identical engine state and outputs.  Then (d) the closed loop over parts (lle_vec_parts_*): a
recorded action stream of min(steps, 4000) steps fed part by part against the same stream fed to plain steps.
python tools/soak.py [steps] [envs]"""
import os, sys, time, zlib
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lle_b200

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536


def digest(v):
    v.synchronize()
    out = []
    raw = v.export_raw()
    for name in ("obs", "state", "avail", "reward", "done", "events", "actions"):
        out.append(zlib.crc32(getattr(v, name).cpu().numpy().tobytes()))
    for k in sorted(raw):
        out.append(zlib.crc32(raw[k].cpu().numpy().tobytes()))
    return out


def run(mode):
    v = lle_b200.VecWorld(lle_b200.Map(level=6), envs, seed=11)
    t0 = time.time()
    done = 0
    while done < steps:
        if mode == "single":
            v.step(None); done += 1
        elif mode == "rollout128":
            k = min(128, steps - done); v.rollout(k); done += k
        else:
            k = min(7, steps - done); v.rollout(k); done += k
            if done < steps:
                v.step(None); done += 1
    d = digest(v)
    return d, time.time() - t0


ref = None
for mode in ("single", "rollout128", "mixed"):
    d, secs = run(mode)
    print(mode, f"{steps} steps x {envs} envs in {secs:.1f} s", d[:4])
    if ref is None:
        ref = d
    assert d == ref, f"{mode} differs from single-step execution"
print("soak ok: identical state and outputs in all three modes")


# (d) the parts loop against plain stepping on the same supplied actions
K = min(steps, 4000)
src = lle_b200.VecWorld(lle_b200.Map(level=6), envs, seed=12)
A = src.n_agents
rec = torch.empty((K, envs, A), dtype=torch.int8).pin_memory()
for t in range(K):
    src.step(None)
    rec[t].copy_(src.actions, non_blocking=True)
d_plain = digest(src)
del src
rec_np = rec.numpy()
v = lle_b200.VecWorld(lle_b200.Map(level=6), envs, seed=12)
act = torch.empty((envs, A), dtype=torch.int8).pin_memory()
rew = torch.empty((envs, 1), dtype=torch.float32).pin_memory()
done = torch.empty((envs,), dtype=torch.uint8).pin_memory()
act_np = act.numpy()
t0 = time.time()
with v.parts_loop(8, act, rew, done) as loop:
    sl = [loop.slice(k) for k in range(loop.n_parts)]
    loop.launch()
    for k in range(loop.n_parts):
        act_np[sl[k]] = rec_np[0][sl[k]]
        loop.feed(k)
    loop.launch()
    for t in range(K):
        for k in range(loop.n_parts):
            loop.wait(k)
            if t + 1 < K:
                act_np[sl[k]] = rec_np[t + 1][sl[k]]
                loop.feed(k)
        if t + 2 < K:
            loop.launch()
secs = time.time() - t0
v.step_count = K
d_parts = digest(v)
print("parts loop", f"{K} steps x {envs} envs in {secs:.2f} s ({envs * K / secs:.3e} env-steps/s from Python)", d_parts[:4])
assert d_parts == d_plain, "the parts loop differs from plain stepping on the same actions"
print("soak ok: the parts loop ends in the same state and outputs")
