"""The oracle against the committed golden fixture (regression anchor for the checker itself)."""
import os

import numpy as np

from _util import GOLDEN, level_text, rollout_digest, synthetic_map
from oracle import lle_oracle as lo


def test_oracle_reproduces_rollout_fixture():
    fx = np.load(os.path.join(GOLDEN, "rollout_lvl6.npz"))
    ora = lo.OracleVec([level_text(6)], None, int(fx["n_envs"]), seed=int(fx["seed"]))
    for t in range(int(fx["steps"])):
        ora.step(None, n_threads=2)
        assert rollout_digest(ora) == fx["digests"][t].tolist(), f"step {t}"
    assert np.array_equal(ora.obs, fx["final_obs"])
    assert np.array_equal(ora.state, fx["final_state"])


def test_synthetic_map_parses_and_has_the_config5_shape():
    text = synthetic_map(64, 64, 8, 16, seed=5)
    w = lo.World(text)
    assert (w.height, w.width, w.n_agents, w.n_sources) == (64, 64, 8, 16)
    assert max(s.beam_len for s in w.laser_sources) <= 63
    assert w.observe_layered().shape == (8, 20, 64, 64)
