"""Quick device-timed loop of the step kernel (development aid, not the contract bench)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lle_b200

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=6)
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--no-obs", action="store_true")
ap.add_argument("--map", type=str, default=None)
ap.add_argument("--rollout", type=int, default=0)
ap.add_argument("--obs-type", type=str, default="layered")
ap.add_argument("--sync", action="store_true", help="synchronise after every step (isolated launches)")
args = ap.parse_args()
moe = None
if args.map == "generated":
    import json
    import numpy as np
    texts = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "generated_5x5.json")))["maps"]
    maps = [lle_b200.Map(t) for t in texts]
    moe = np.repeat(np.arange(len(maps), dtype=np.int32), args.envs // len(maps))
elif args.map == "synthetic":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from _util import synthetic_map
    maps = lle_b200.Map(synthetic_map(64, 64, 8, 16, seed=5))
else:
    maps = lle_b200.Map(level=args.level)
vec = lle_b200.VecWorld(maps, args.envs, seed=1, write_obs=not args.no_obs, map_of_env=moe, obs_type=args.obs_type)
for _ in range(50):
    vec.step(None)
vec.synchronize()
vec.timing_begin()
if args.rollout:
    for _ in range(args.steps // args.rollout):
        vec.rollout(args.rollout)
    ms, n = vec.timing_end()
    n = (args.steps // args.rollout) * args.rollout
else:
    for _ in range(args.steps):
        vec.step(None)
        if args.sync:
            vec.synchronize()
    ms, n = vec.timing_end()
us = ms * 1e3 / n
obs_bytes = (vec.obs[0].numel() if vec.obs is not None else 0) * 4 * args.envs
print(json.dumps({"level": args.map or args.level, "obs_type": args.obs_type, "envs": args.envs, "obs": not args.no_obs, "rollout": args.rollout, "us_per_step": round(us, 2),
                  "env_steps_per_s": round(args.envs / (us * 1e-6)), "obs_GBps": round(obs_bytes / (us * 1e-6) / 1e9, 1)}))
