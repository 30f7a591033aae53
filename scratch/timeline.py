import ctypes as C, os, sys
os.environ["LLE_B200_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, lle_b200
from lle_b200._native import lib
level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
roll = int(sys.argv[2]) if len(sys.argv) > 2 else 0
vec = lle_b200.VecWorld(lle_b200.Map(level=level), 65536, seed=1)
for _ in range(20): vec.step(None)
if roll: vec.rollout(roll)
else: vec.step(None)
n = C.c_int64(0)
lib().lle_vec_debug_timeline(vec._h, None, 0, C.byref(n))
buf = np.zeros((n.value, 4), dtype=np.uint64)
lib().lle_vec_debug_timeline(vec._h, buf.ctypes.data, n.value, C.byref(n))
t = buf.astype(np.int64)
t0 = t[:, 0].min()
rel = (t - t0) / 1e3
act = t[:, 1] > 0
print(f"warps {n.value} active {act.sum()}  start: max {rel[:,0].max():.1f} us")
print(f"first store: min {rel[act,1].min():.1f} median {np.median(rel[act,1]):.1f} max {rel[act,1].max():.1f} us")
print(f"last store : min {rel[act,2].min():.1f} median {np.median(rel[act,2]):.1f} max {rel[act,2].max():.1f} us")
print(f"warp end   : min {rel[:,3].min():.1f} median {np.median(rel[:,3]):.1f} max {rel[:,3].max():.1f} us")
h, e = np.histogram(rel[:, 3], bins=10)
print("end histogram", list(zip(np.round(e[:-1], 1), h)))
