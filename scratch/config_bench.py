"""Device-timed throughput + HBM roofline of the other BASELINE.json configs (bench.py itself measures configs[1]).
One JSON line per config: python scratch/config_bench.py --config 1|3|4|5 [--envs N] [--steps K]
  1: World.level(1) x 65,536            3: 1,024 generated 5x5 maps x 1,024 envs
  4: levels 1-6 mixed, 2,097,152 envs   5: synthetic 64x64, 8 agents, 16 sources x 262,144 envs (86 GB of observations)
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import lle_b200
from bench import algorithmic_bytes, measured_peak

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True)
ap.add_argument("--envs", type=int, default=0)
ap.add_argument("--steps", type=int, default=0)
args = ap.parse_args()
cfg = args.config
defaults = {1: (65536, 4000), 3: (1 << 20, 500), 4: (1 << 21, 200), 5: (262144, 30)}
n_envs = args.envs or defaults[cfg][0]
steps = args.steps or defaults[cfg][1]
if cfg == 1:
    parts, name = [lle_b200.VecWorld(lle_b200.Map(level=1), n_envs, seed=1)], "World.level(1) (BASELINE.json configs[0] map, batched)"
elif cfg == 3:
    texts = json.load(open(os.path.join(ROOT, "tests", "golden", "generated_5x5.json")))["maps"]
    maps = [lle_b200.Map(t) for t in texts]
    moe = np.repeat(np.arange(len(maps), dtype=np.int32), n_envs // len(maps))
    parts, name = [lle_b200.VecWorld(maps, n_envs, map_of_env=moe, seed=1)], "1,024 generated 5x5 maps, 2 agents, 2 lasers (configs[2])"
elif cfg == 4:
    grp = lle_b200.VecWorldGroup([(lle_b200.Map(level=l), n_envs // 6 + (1 if l <= n_envs % 6 else 0)) for l in range(1, 7)], seed=1)
    parts, name = grp.parts, "levels 1-6 mixed, one sub-batch per level (configs[3] per-GPU slice)"
else:
    from _util import synthetic_map
    parts, name = [lle_b200.VecWorld(lle_b200.Map(synthetic_map(64, 64, 8, 16, seed=5)), n_envs, seed=1)], "synthetic 64x64, 8 agents, 16 sources (configs[4])"
for _ in range(20):
    for p in parts:
        p.step(None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    for p in parts:
        p.step(None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
total_envs = sum(p.n_envs for p in parts)
bytes_step = sum(algorithmic_bytes(p.n_agents, p.n_gems, p.n_channels, p.height, p.width, p.reward_dim, p.record_bytes)["total"] * p.n_envs for p in parts)
agent_steps = sum(p.n_envs * p.n_agents for p in parts)
peak, src = measured_peak()
ach = bytes_step / (ms / 1e3) / 1e9
print(json.dumps({"config": cfg, "workload": name, "envs": total_envs, "steps": steps, "ms_per_step": ms, "env_steps_per_s": total_envs / (ms / 1e3),
                  "agent_env_steps_per_s": agent_steps / (ms / 1e3), "algorithmic_bytes_per_step": bytes_step,
                  "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src}}))
