"""One generator launch for ncu (development aid): python scratch/gen_prof.py [n]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lle_b200.generator import WorldGenerator
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 92 * 8
g = WorldGenerator(width=5, height=5, n_agents=2, n_lasers=2, batch=n)
for r in range(3):
    g.run(first_seed=r * n, n=n)
torch.cuda.synchronize()
print("ok", float(g.status[:n].float().mean()))
