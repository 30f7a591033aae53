"""Small case for compute-sanitizer: a few steps / reset / set_state / rollout on tiny batches of three map shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, lle_b200
from _util import level_text, synthetic_map
for maps, n in ((level_text(6), 96), (level_text(1), 40), ("S0 . G\nS1 X X", 70), (synthetic_map(64, 64, 8, 16, seed=5), 16)):
    v = lle_b200.VecWorld(maps, n, seed=3)
    for _ in range(4):
        v.step(None)
    v.rollout(3)
    v.reset()
    acts = torch.full((n, v.n_agents), 4, dtype=torch.int8)
    v.step(acts)
    pos = v.export_raw()["pos"].to(torch.int32)
    v.set_state(pos, torch.zeros((n, v.n_gems), dtype=torch.uint8), torch.ones((n, v.n_agents), dtype=torch.uint8))
    v.synchronize()
    print("ok", n, v.n_agents, int(v.err.sum()))
