"""Quick tour of lle_b200 on one GPU: the reference's single-world API, the batched environment, device-generated maps.
Run: python examples/quickstart.py [n_envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lle_b200 as lle  # noqa: E402  (the import paths of the reference's `lle` package are kept)
from lle_b200.generator import generate  # noqa: E402

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096

# 1. lle.World, as in the reference (python/lle/world): one world, N = 1 on the device
world = lle.World.level(6)
world.reset()
events = world.step([lle.Action.SOUTH] * world.n_agents)
print("World.level(6):", world.agents_positions, [e.event_type.name for e in events], "joint actions:", len(world.available_joint_actions()))

# 2. the batched environment: every per-step product is a device tensor (zero copy; .__dlpack__() exports it)
env = lle.level(6).n_envs(n_envs).seed(0).obs_type(lle.ObservationType.LAYERED).build()
obs, state = env.reset()
episodes, total = 0, 0.0
for _ in range(200):
    obs, state, reward, done = env.step()          # device-sampled actions; or env.step(actions_i8_cuda) from a policy
    episodes += int(done.sum())
    total += float(reward.sum())
torch.cuda.synchronize()
print(f"VecLLE: obs {tuple(obs.shape)} {obs.dtype} on {obs.device}; 200 steps x {n_envs} envs, {episodes} episodes, reward sum {total:.0f}")

# 2b. a policy that runs on the HOST between steps (closed loop): one launch per step, the kernel waits part by part for the actions
rng_state = {"episodes": 0}


def host_policy(part, reward, done):                # numpy views of this part's results in pinned host memory
    rng_state["episodes"] += int(done.sum())
    return None                                     # None: everybody STAYs; or an int8 array (len(part), n_agents)


env.run_host_policy(host_policy, steps=50, n_parts=8)
print(f"host policy in the loop: 50 steps, {rng_state['episodes']} episodes seen by the policy")

# 3. lle.generate(...): layouts generated on the device (bit-identical per seed to the reference's Python generator), then stepped
maps = list(generate(5, 5, 2).lasers(2).walls(2).needs_blocker().take(64, seed=0, distinct=True, texts=True))
batch = lle.VecWorld(maps, 64 * 16, map_of_env=[m for m in range(64) for _ in range(16)], seed=1)
batch.rollout(100)                                  # 100 lock-step steps in one launch
torch.cuda.synchronize()
print(f"generated maps: {len(maps)} distinct 5x5 layouts that need a blocker; first one:\n{maps[0]}\nrollout done flags: {int(batch.done.sum())}")
print("ok")
