"""One lle_vec_rollout(128) launch of the headline workload after a few single steps (subject of the ncu capture in
profiles/capture.sh: launch 7 of lle_world_kernel = the reset at creation, 5 steps, then the rollout)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lle_b200

vec = lle_b200.VecWorld(lle_b200.Map(level=6), 65536, seed=2026)
for _ in range(5):
    vec.step(None)
torch.cuda.synchronize()
vec.rollout(128)
torch.cuda.synchronize()
print("steps", vec.step_count)
