"""Generates tests/golden/rollout_lvl6.npz from the ORACLE: per-step checksums of every output of a seeded
Philox rollout on level 6 plus the final observation/state.  The CUDA path must reproduce it bit for bit
(tests/test_gpu_parity.py::test_golden_fixture_level6); tests/test_golden_cpu.py checks that the oracle still does.

The reference itself cannot run in this image (Rust toolchain absent), so this fixture pins the oracle's
behaviour at the commit where it passed the transcribed reference KATs; it is a regression anchor, not an
independent source of truth.   Usage: python tests/golden/make_rollout_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
from _util import level_text, rollout_digest  # noqa: E402
from oracle import lle_oracle as lo  # noqa: E402

N, STEPS, SEED = 192, 160, 20261018
ora = lo.OracleVec([level_text(6)], None, N, seed=SEED)
digests = []
for t in range(STEPS):
    ora.step(None)
    digests.append(rollout_digest(ora))
np.savez_compressed(os.path.join(HERE, "rollout_lvl6.npz"), n_envs=N, steps=STEPS, seed=SEED,
                    digests=np.array(digests, dtype=np.uint64), final_obs=ora.obs.copy(), final_state=ora.state.copy())
print("written", len(digests), "steps")
