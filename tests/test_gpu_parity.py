"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the oracle
on the same seeded inputs, bit for bit — outputs (obs, state, avail, reward, done, events, actions, err) and
the raw engine state (positions, alive/arrived, tile slots, beam masks, gems)."""
import os

import numpy as np
import pytest
import torch

from _parity import assert_same
from _util import level_text
from oracle import lle_oracle as lo

pytestmark = pytest.mark.gpu


class Dev:
    """Adapter: host copies of a VecWorld's buffers with the oracle's attribute names."""

    def __init__(self, vec):
        self.vec = vec

    def pull(self):
        v = self.vec
        v.synchronize()
        for name in ("obs", "state", "avail", "reward", "done", "events", "actions", "err"):
            setattr(self, name, getattr(v, name).cpu().numpy())
        raw = {k: t.cpu().numpy() for k, t in v.export_raw().items()}
        raw["beam_on"] = raw["beam_on"].view(np.uint64)
        raw["collected"] = raw["collected"].view(np.uint64)
        if getattr(self, "with_extras", False):
            raw["extras"] = v.extras.cpu().numpy()
        return raw


def make_pair(maps, map_of_env, n_envs, **kw):
    import lle_b200

    okw = dict(multi_objective=kw.get("reward_dim", 1) == 4, walkable_lasers=kw.get("walkable_lasers", True),
               auto_reset=kw.get("auto_reset", True), seed=kw.get("seed", 0), env_id_base=kw.get("env_id_base", 0),
               extras=kw.get("extras"), pbrs=kw.get("pbrs"), obs_type=kw.get("obs_type", "layered"),
               padding_size=kw.get("padding_size", 0), randomize_lasers=kw.get("randomize_lasers", False))
    ora = lo.OracleVec(maps, map_of_env, n_envs, **okw)
    vec = lle_b200.VecWorld(maps, n_envs, map_of_env=map_of_env, **kw)
    dev = Dev(vec)
    dev.with_extras = ora.JE > 0
    return ora, dev


def run_pair(maps, map_of_env, n_envs, steps, check_every=1, **kw):
    ora, dev = make_pair(maps, map_of_env, n_envs, **kw)
    assert_same(dev, ora, dev.pull(), "after reset")
    for t in range(steps):
        ora.step(None)
        dev.vec.step(None)
        if t % check_every == 0 or t == steps - 1:
            assert_same(dev, ora, dev.pull(), f"step {t}")
    return ora, dev


def test_levels_philox_rollout():
    for level in range(1, 7):
        run_pair([level_text(level)], None, 200, 200, seed=level)


def test_layout_corpus_philox_rollout(layouts):
    for k, (name, text) in enumerate(sorted(layouts.items())):
        run_pair([text], None, 96, 100, seed=100 + k)


def test_config1_anchor_level1():
    """BASELINE config 1: lvl1, N=4096 envs, device Philox actions replayed in the oracle."""
    run_pair([level_text(1)], None, 4096, 300, check_every=7, seed=1234)


def test_config1_anchor_full_length():
    """SURVEY 8d config 1 as specified: T = 10,000 steps x N = 4,096 envs of level 1, compared bit-exactly at 100+ points
    of the rollout (every output and the raw engine state)."""
    run_pair([level_text(1)], None, 4096, 10000, check_every=89, seed=4321)


def test_config2_level6_full_size():
    """BASELINE config 2 at full size: lvl6, 65,536 envs, layered observations."""
    run_pair([level_text(6)], None, 65536, 12, seed=77)


def test_ragged_batch_sizes():
    for n in (1, 31, 33, 100):
        run_pair([level_text(6)], None, n, 40, seed=n)


def test_heterogeneous_maps_in_one_batch():
    maps = [level_text(2), level_text(3), level_text(4)]
    moe = [(e * 7 + e // 5) % 3 for e in range(1000)]
    run_pair(maps, moe, 1000, 150, seed=11, env_id_base=1000)


def test_options():
    run_pair([level_text(6)], None, 256, 150, reward_dim=4, seed=7)
    run_pair([level_text(5)], None, 256, 150, walkable_lasers=False, seed=8)
    ora, dev = run_pair([level_text(6)], None, 256, 250, auto_reset=False, seed=5)
    assert (ora.err == 2).any()


def test_laser_subgoal_extras_and_pbrs():
    """LaserSubgoal flags (extras_generators.py:75-101) and PotentialShapedLLE (reward_strategy.py:113-181)
    fused into the step, single- and multi-objective, all / selected sources, with and without auto-reset."""
    run_pair([level_text(6)], None, 300, 200, extras="laser_subgoal", seed=31)
    run_pair([level_text(5)], None, 300, 200, pbrs=dict(gamma=0.99, reward_value=0.5), seed=32)
    run_pair([level_text(6)], None, 300, 200, reward_dim=4, pbrs=dict(gamma=0.9, reward_value=0.3, lasers_to_reward=[2, 0]), seed=33)
    run_pair([level_text(4)], None, 300, 200, pbrs=dict(with_extras=False, lasers_to_reward=[1]), extras=[0], seed=34)
    run_pair([level_text(6)], None, 200, 200, auto_reset=False, pbrs=dict(gamma=1.0, reward_value=1.0), seed=35)
    rows = [" .   .   . . . ."] + [f"S{k}  L{k}W  . . . X" for k in range(14)]
    run_pair(["\n".join(rows)], None, 64, 80, pbrs=dict(), seed=36)  # 14 sources: extras_dim 14
    from _util import synthetic_map

    run_pair([synthetic_map(64, 64, 8, 16, seed=5)], None, 64, 30, check_every=3, pbrs=dict(), seed=37)


@pytest.mark.parametrize("obs_type", ["partial3x3", "partial5x5", "partial7x7", "perspective", "layered-padded-2", "state",
                                      "normalized-state"])
def test_observation_types(obs_type):
    """the other observation generators (observations.py:141-158, :196-214, :296-395) as store epilogues of
    the same fused step, bit-exact against the oracle's restatement under Philox rollouts."""
    for level in (3, 5, 6):
        run_pair([level_text(level)], None, 200, 120, obs_type=obs_type, seed=50 + level)
    # heterogeneous maps in one batch, ragged size, no auto-reset (dead agents stay in the observation)
    maps = [level_text(2), level_text(3), level_text(4)]
    moe = [(e * 7 + e // 5) % 3 for e in range(333)]
    run_pair(maps, moe, 333, 100, obs_type=obs_type, auto_reset=False, seed=57)
    run_pair(["S0 . G\nS1 X X"], None, 70, 40, obs_type=obs_type, seed=58)  # tiny map: windows mostly off the map


@pytest.mark.parametrize("obs_type", ["partial3x3", "partial5x5", "partial7x7"])
def test_partial_observations_on_the_general_kernel(obs_type, monkeypatch):
    """Worlds with a small record take the thread-per-world kernel for partial observations too (tiny_kernel.cuh PARTIAL, covered by
    test_observation_types); the warp-per-group renderer of world_kernel.cuh stays in use for everything else: same cases on it."""
    monkeypatch.setenv("LLE_B200_NO_TINY", "1")
    for level in (3, 6):
        run_pair([level_text(level)], None, 200, 100, obs_type=obs_type, seed=150 + level)
    maps = [level_text(2), level_text(3), level_text(4)]
    moe = [(e * 7 + e // 5) % 3 for e in range(333)]
    run_pair(maps, moe, 333, 80, obs_type=obs_type, auto_reset=False, seed=157)


@pytest.mark.parametrize("obs_type", ["partial5x5", "partial7x7"])
def test_large_windows_forced_onto_the_thread_per_world_kernel(obs_type, monkeypatch):
    """By default only windows under 2 KB per world (3x3 on level 6) take tiny_kernel.cuh's PARTIAL instantiation; the kernel itself
    handles any odd size (several rounds per ticket, one world per tile for 7x7)."""
    monkeypatch.setenv("LLE_B200_TINY_PARTIAL", "1")
    for level in (3, 6):
        run_pair([level_text(level)], None, 200, 100, obs_type=obs_type, seed=160 + level)
    run_pair(["S0 . G\nS1 X X"], None, 70, 40, obs_type=obs_type, seed=168)


def test_observation_types_large_and_many_agents():
    from _util import synthetic_map

    big = synthetic_map(64, 64, 8, 16, seed=5)
    run_pair([big], None, 64, 30, check_every=3, obs_type="perspective", seed=61)   # 8 x 327,680 B per world, chunked
    run_pair([big], None, 64, 30, check_every=3, obs_type="partial7x7", seed=62)    # 8 x 19 x 49 floats per world
    run_pair([big], None, 64, 30, check_every=3, obs_type="layered-padded", padding_size=4, seed=63)
    rows = [" .   .   . . . ."] + [f"S{k}  L{k}W  . . . X" for k in range(14)]
    run_pair(["\n".join(rows)], None, 64, 40, obs_type="partial3x3", seed=64)      # 14 agents
    run_pair(["\n".join(rows)], None, 64, 40, obs_type="perspective", seed=65)
    run_pair(["S0 L1E X"], None, 40, 10, obs_type="partial3x3", seed=66)            # colour == n_agents spills into GEM
    import lle_b200

    with pytest.raises(IndexError):  # colour 3 >= n_agents + 2: numpy IndexError in the reference
        lle_b200.VecLLE(["S0 L3E X"], 4, obs_type="partial3x3")
    v = lle_b200.VecWorld(level_text(6), 16, obs_type="flattened")
    assert tuple(v.obs.shape) == (16, 12 * 12 * 13) and tuple(v.obs_per_agent.shape) == (16, 4, 12 * 12 * 13)


def test_laser_source_mutators_in_a_batch():
    """lle_vec_set_source (LaserBeam::set_agent_id / enable / disable, laser.rs:69-84) on one map of a heterogeneous batch,
    then Philox rollouts: every output and the raw engine state stay bit-exact against the oracle."""
    maps = [level_text(4), level_text(3), level_text(4)]  # 2 agents, 1 gem; sources (colour): lvl4 (0), (1); lvl3 (0)
    moe = [e % 3 for e in range(300)]
    for kw in (dict(), dict(obs_type="partial5x5"), dict(walkable_lasers=False), dict(obs_type="perspective", auto_reset=False)):
        ora, dev = make_pair(maps, moe, 300, seed=70, **kw)
        for t in range(30):
            ora.step(None); dev.vec.step(None)
        for target in (ora, dev.vec):
            target.set_source(1, enabled=False, map_index=0)   # level 4, second source, in the envs of map 0 only
            target.set_source(0, agent_id=1, map_index=1)      # level 3: recoloured
            target.set_source(0, agent_id=1, map_index=2)      # level 4 (third map): both sources now colour 1
        assert dev.vec.source_states(0) == [(0, True), (1, False)] and dev.vec.source_states(2) == [(1, True), (1, True)]
        raw = dev.pull()
        assert np.array_equal(raw["beam_on"][:, :ora.NB], np.asarray(ora.beam_on)[:, :ora.NB])
        ora.reset(); dev.vec.reset()  # LLE.reset refreshes the cached static layers (observations.py:128-137)
        assert_same(dev, ora, dev.pull(), "after the reset that follows the mutation")
        for t in range(120):
            ora.step(None); dev.vec.step(None)
            if t % 4 == 0:
                assert_same(dev, ora, dev.pull(), f"step {t} after the mutation")
        for target in (ora, dev.vec):
            target.set_source(1, enabled=True, map_index=0)    # the whole beam comes back on, whoever stands in it
            target.set_exits([(11, 0), (10, 3), (5, 5)], map_index=1)  # World::set_exit_positions (world.rs:195-234)
        raw = dev.pull()
        assert np.array_equal(raw["beam_on"][:, :ora.NB], np.asarray(ora.beam_on)[:, :ora.NB])
        for t in range(60):  # exits moved mid-episode: the reference's cached EXIT layer only refreshes at each env's next
            ora.step(None); dev.vec.step(None)  # reset (observations.py:128-137), so observations are not compared here
        raw = dev.pull()
        for name in ("pos", "alive", "arrived", "slot", "collected"):
            assert np.array_equal(raw[name], np.asarray(getattr(ora, name))), name
        ora.reset(); dev.vec.reset()
        for t in range(80):
            ora.step(None); dev.vec.step(None)
            assert_same(dev, ora, dev.pull(), f"step {t} after re-enabling and moving the exits")


def test_random_maps_fuzz():
    """Differential fuzzing on the device: 400 random maps (crossing beams, foreign colours, voids, starts next to beams),
    each stepped under Philox actions with a rotating observation type / reward / auto-reset setting."""
    import random

    from _util import random_map

    rng = random.Random(777)
    settings = [dict(), dict(obs_type="partial3x3"), dict(obs_type="perspective"), dict(reward_dim=4), dict(auto_reset=False),
                dict(walkable_lasers=False), dict(obs_type="partial5x5", auto_reset=False), dict(pbrs=dict()),
                dict(obs_type="layered-padded-1"), dict(obs_type="normalized-state", extras="laser_subgoal")]
    n_run = 0
    while n_run < 400:
        text = random_map(rng)
        try:
            lo.World(text)
        except lo.ParsingError:
            continue
        kw = dict(settings[n_run % len(settings)])
        try:
            ora, dev = make_pair([text], None, 48, seed=1000 + n_run, **kw)
        except IndexError:  # a foreign colour selects a channel past the last layer: both sides must refuse
            import lle_b200
            with pytest.raises(IndexError):
                lle_b200.VecLLE([text], 4, **{k: v for k, v in kw.items() if k in ("obs_type",)})
            n_run += 1
            continue
        for t in range(40):
            ora.step(None)
            dev.vec.step(None)
            if t % 3 == 0 or t == 39:
                assert_same(dev, ora, dev.pull(), f"map {n_run} ({kw}) step {t}\n{text}")
        n_run += 1


def test_randomize_lasers():
    """LLE(randomize_lasers=True) (env.py:198-200): every LLE-level reset (explicit, masked, automatic) recolours the env's
    sources; the colours come from the library's own Philox stream, restated by the oracle (Python's `random` is unpinned)."""
    for maps, moe, n, kw in (([level_text(6)], None, 400, dict()), ([level_text(4)], None, 300, dict(obs_type="partial5x5", walkable_lasers=False)),
                             ([level_text(3), level_text(4)], [e % 2 for e in range(256)], 256, dict(reward_dim=4, obs_type="perspective")),
                             ([level_text(5)], None, 200, dict(auto_reset=False, extras="laser_subgoal"))):
        ora, dev = make_pair(maps, moe, n, seed=95, randomize_lasers=True, **kw)
        assert_same(dev, ora, dev.pull(), "construction: the colours of the text")
        colours0 = dev.vec.source_colours().cpu().numpy()
        for t in range(150):
            ora.step(None); dev.vec.step(None)
            if t % 5 == 0:
                assert_same(dev, ora, dev.pull(), f"step {t} ({kw})")
        ora.reset(); dev.vec.reset()
        assert_same(dev, ora, dev.pull(), "explicit reset")
        colours1 = dev.vec.source_colours().cpu().numpy()
        assert (colours1 != colours0).any() and colours1.min() >= 0 and colours1.max() < ora.A
        if ora.NB >= 2 and n >= 256:
            assert len({tuple(c) for c in colours1}) > 2  # envs draw independently
        mask = (np.arange(n) % 3 == 0).astype(np.uint8)
        ora.reset(mask); dev.vec.reset(torch.from_numpy(mask).cuda())
        assert_same(dev, ora, dev.pull(), "masked reset")
        colours2 = dev.vec.source_colours().cpu().numpy()
        assert np.array_equal(colours2[mask == 0], colours1[mask == 0])
        for t in range(60):
            ora.step(None); dev.vec.step(None)
        assert_same(dev, ora, dev.pull(), "after more steps")
    import lle_b200
    with pytest.raises(ValueError):  # a laser across a start position: set_colour would raise in the reference
        lle_b200.VecWorld("S0 . X\nL0E S1 X", 4, randomize_lasers=True)


def test_masked_reset():
    """lle_vec_reset with a per-env mask: only the flagged envs start a new episode (and have their transition outputs
    cleared); the others keep everything."""
    rng = np.random.default_rng(5)
    for kw in (dict(), dict(auto_reset=False, reward_dim=4), dict(obs_type="partial3x3", extras="laser_subgoal", pbrs=dict())):
        ora, dev = make_pair([level_text(6)], None, 500, seed=90, **kw)
        for rnd in range(6):
            for _ in range(15):
                ora.step(None); dev.vec.step(None)
            mask = (rng.random(500) < (0.0 if rnd == 3 else 0.4)).astype(np.uint8)
            ora.reset(mask); dev.vec.reset(torch.from_numpy(mask).cuda())
            assert_same(dev, ora, dev.pull(), f"masked reset {rnd} ({kw})")
        ora.step(None); dev.vec.step(None)
        assert_same(dev, ora, dev.pull(), "step after the masked resets")


def test_mixed_operation_sequences():
    """A random interleaving of every way to drive a vec — single steps (sampled / supplied), rollouts, the host pipeline,
    full and masked resets, set_state, refresh, source and exit mutators — checked against the oracle after every
    operation.  Exercises the ordering between launch kinds (programmatic dependent launches, narrow grids, epoch flags)."""
    import ctypes as C

    rng = np.random.default_rng(2026)
    L = lo.lib()
    n = 777
    ora, dev = make_pair([level_text(6)], None, n, seed=91)
    vec = dev.vec
    A, R = ora.A, ora.R
    pinned = [(torch.empty((n, A), dtype=torch.int8).pin_memory(), torch.empty((n, R), dtype=torch.float32).pin_memory(),
               torch.empty((n,), dtype=torch.uint8).pin_memory()) for _ in range(4)]
    for op_index in range(160):
        op = rng.choice(["steps", "supplied", "rollout", "pipeline", "reset", "masked", "set_state", "refresh", "source", "exits"],
                        p=[0.25, 0.1, 0.15, 0.15, 0.05, 0.08, 0.07, 0.05, 0.05, 0.05])
        if op == "steps":
            for _ in range(int(rng.integers(1, 12))):  # back to back, no synchronisation in between
                ora.step(None); vec.step(None)
        elif op == "supplied":
            ora.step(None)  # the oracle samples a valid joint action; the device replays it (some made invalid)
            vec.step(torch.from_numpy(np.array(ora.actions)).cuda())
        elif op == "rollout":
            k = int(rng.integers(1, 9))
            vec.rollout(k)
            for _ in range(k):
                ora.step(None)
        elif op == "pipeline":
            k = int(rng.integers(1, 5))
            expect = []
            for q in range(k):
                ora.step(None)
                expect.append((np.array(ora.reward), np.array(ora.done)))
                pinned[q][0].copy_(torch.from_numpy(np.array(ora.actions)))
                vec.submit_host(pinned[q][0], pinned[q][1], pinned[q][2])
            for q in range(k):
                vec.wait_host()
                assert np.array_equal(pinned[q][1].numpy(), expect[q][0]) and np.array_equal(pinned[q][2].numpy(), expect[q][1])
        elif op == "reset":
            ora.reset(); vec.reset()
        elif op == "masked":
            mask = (rng.random(n) < 0.3).astype(np.uint8)
            ora.reset(mask); vec.reset(torch.from_numpy(mask).cuda())
        elif op == "set_state":
            pos = np.stack([rng.integers(0, ora.H, size=(n, A)), rng.integers(0, ora.W, size=(n, A))], axis=-1).astype(np.int32)
            gems = rng.integers(0, 2, size=(n, ora.G)).astype(np.uint8)
            alive = (rng.random((n, A)) < 0.9).astype(np.uint8)
            vec.set_state(torch.from_numpy(pos), torch.from_numpy(gems), torch.from_numpy(alive))
            for e in range(n):
                p = (C.c_long * (2 * A))(*[int(x) for x in pos[e].reshape(-1)])
                g = (C.c_uint8 * max(1, ora.G))(*[int(x) for x in gems[e]])
                a = (C.c_uint8 * A)(*[int(x) for x in alive[e]])
                L.lleo_vec_set_state_env(ora._h, C.c_long(e), p, A, g, ora.G, a)
            L.lleo_vec_refresh(ora._h)
            raw = dev.pull()
            for name in ("pos", "alive", "arrived", "slot", "collected"):
                assert np.array_equal(raw[name], np.asarray(getattr(ora, name))), f"op {op_index} set_state raw {name}"
            assert np.array_equal(raw["beam_on"][:, :ora.NB], np.asarray(ora.beam_on)[:, :ora.NB])
            ora.reset(); vec.reset()  # err / event bytes of a failed set_state are compared in test_set_state_fuzz
        elif op == "refresh":
            vec.refresh()
        elif op == "source":
            b, colour, enabled = int(rng.integers(0, 3)), int(rng.integers(0, 4)), bool(rng.integers(0, 2))
            for target in (ora, vec):
                target.set_source(b, agent_id=colour, enabled=enabled)
            ora.reset(); vec.reset()  # the reference's cached static layers follow at the next reset
        elif op == "exits":
            exits = [(11, int(j)) for j in rng.choice(12, size=5, replace=False)]  # (11, 12) holds a gem
            for target in (ora, vec):
                target.set_exits(exits)
            ora.reset(); vec.reset()
        assert vec.step_count == ora.t, (op, vec.step_count, ora.t)
        assert_same(dev, ora, dev.pull(), f"op {op_index}: {op}")


def test_supplied_actions_with_invalid_ones():
    rng = np.random.default_rng(0)
    ora, dev = make_pair([level_text(5)], None, 512, seed=1)
    n_bad = 0
    for t in range(100):
        actions = rng.integers(0, 5, size=(512, ora.A)).astype(np.int8)
        if t % 10 == 0:
            actions[::17, 0] = 7
        ora.step(actions)
        dev.vec.step(torch.from_numpy(actions).cuda())
        assert_same(dev, ora, dev.pull(), f"step {t}")
        n_bad += int((ora.err == 1).sum())
    assert n_bad > 100


def test_pipelined_host_stepping():
    """lle_vec_pipeline_submit / _wait: host actions in, reward + done out, up to 8 steps in flight; every step's host
    results and the final device buffers equal the oracle's."""
    for depth, n, level in ((1, 100, 5), (2, 1000, 6), (4, 333, 6), (8, 4096, 6)):
        ora, dev = make_pair([level_text(level)], None, n, seed=40 + depth)
        vec = dev.vec
        steps = 60
        slots = [(torch.empty((n, ora.A), dtype=torch.int8).pin_memory(), torch.empty((n, ora.R), dtype=torch.float32).pin_memory(),
                  torch.empty((n,), dtype=torch.uint8).pin_memory()) for _ in range(depth)]
        expect = []
        inflight = []

        def retire():
            k = inflight.pop(0)
            vec.wait_host()
            _, rew, done = slots[k % depth]
            assert np.array_equal(rew.numpy(), expect[k][0]), f"depth {depth}: reward of step {k}"
            assert np.array_equal(done.numpy(), expect[k][1]), f"depth {depth}: done of step {k}"

        for t in range(steps):
            if len(inflight) == depth:
                retire()
            ora.step(None)  # the oracle samples; the device replays the recorded actions from the host
            expect.append((np.array(ora.reward), np.array(ora.done)))
            act, rew, done = slots[t % depth]
            act.copy_(torch.from_numpy(np.array(ora.actions)))
            vec.submit_host(act, rew, done)
            inflight.append(t)
        while inflight:
            retire()
        vec.step_count = ora.t
        assert_same(dev, ora, dev.pull(), f"pipelined depth {depth}")
        # device sampling through the pipeline, and back to plain stepping afterwards
        ora.step(None)
        vec.submit_host(None, slots[0][1], slots[0][2])
        assert vec.wait_host() == 0
        assert np.array_equal(slots[0][1].numpy(), ora.reward)
        ora.step(None)
        vec.step(None)
        assert_same(dev, ora, dev.pull(), f"after the pipeline, depth {depth}")
    with pytest.raises(ValueError):
        vec.submit_host(None, slots[0][1], slots[0][2])
        vec.step(None)  # not drained
    vec.wait_host()


def _parts_rollout(ora, dev, n_parts, steps, ahead=2):
    """Drive `steps` steps through lle_vec_parts_*: the oracle samples, the device replays its actions part by part; the host
    results of every part and step are compared as they arrive."""
    vec = dev.vec
    n = vec.n_envs
    act = torch.empty((n, ora.A), dtype=torch.int8).pin_memory()
    rew = torch.empty((n, ora.R), dtype=torch.float32).pin_memory()
    done = torch.empty((n,), dtype=torch.uint8).pin_memory()
    recorded = []
    for t in range(steps):
        ora.step(None)
        recorded.append((np.array(ora.actions), np.array(ora.reward), np.array(ora.done)))
    with vec.parts_loop(n_parts, act, rew, done) as loop:
        assert 1 <= loop.n_parts <= n_parts and sum(c for _, c in loop.ranges) == n and loop.ranges[0][0] == 0
        for k in range(1, loop.n_parts):
            assert loop.ranges[k][0] == loop.ranges[k - 1][0] + loop.ranges[k - 1][1]
        loop.launch()
        for k in range(loop.n_parts):
            act[loop.slice(k)] = torch.from_numpy(recorded[0][0][loop.slice(k)])
            loop.feed(k)
        for _ in range(1, min(ahead, steps)):
            loop.launch()
        for t in range(steps):
            for k in range(loop.n_parts):
                sl = loop.slice(k)
                loop.wait(k)
                assert np.array_equal(rew.numpy()[sl], recorded[t][1][sl]), f"reward of part {k}, step {t}"
                assert np.array_equal(done.numpy()[sl], recorded[t][2][sl]), f"done of part {k}, step {t}"
                if t + 1 < steps:
                    act[sl] = torch.from_numpy(recorded[t + 1][0][sl])
                    loop.feed(k)
            if t + ahead < steps:
                loop.launch()
    vec.step_count = ora.t
    return act, rew, done


@pytest.mark.parametrize("n_parts,n,level,ahead", [(1, 100, 5, 1), (3, 1000, 6, 2), (8, 4096, 6, 2), (5, 333, 3, 3), (16, 2100, 6, 4),
                                                   (50, 1000, 6, 2), (13, 100, 6, 2), (125, 1000, 6, 3)])  # the last three: padding tickets behind N
def test_parts_loop_equals_the_oracle(n_parts, n, level, ahead):
    """lle_vec_parts_*: one launch per step of the whole batch, the host feeding actions and reading reward / done part by part;
    every part's host results at every step and the final device buffers equal the oracle's."""
    ora, dev = make_pair([level_text(level)], None, n, seed=140 + n_parts)
    _parts_rollout(ora, dev, n_parts, 50, ahead)
    assert_same(dev, ora, dev.pull(), f"after a parts loop of {n_parts} parts")
    ora.step(None)
    dev.vec.step(None)  # plain stepping works again once the loop is closed
    assert_same(dev, ora, dev.pull(), "after the loop")
    _parts_rollout(ora, dev, n_parts, 7, ahead)  # and a second loop on the same vec
    assert_same(dev, ora, dev.pull(), "after a second loop")


def test_parts_loop_other_kernels_and_options():
    """The general kernel (several maps), a vec whose plain steps run on the thread-per-world kernel, multi-objective rewards,
    partial observations, no auto-reset."""
    maps = [level_text(2), level_text(3), level_text(4)]
    moe = [(e * 7 + e // 5) % 3 for e in range(700)]
    ora, dev = make_pair(maps, moe, 700, seed=151, reward_dim=4)
    _parts_rollout(ora, dev, 4, 40)
    assert_same(dev, ora, dev.pull(), "heterogeneous batch")
    tiny = ["S0 . G . X\n. @ . . .\nL1E . . . .\n. . . @ .\nS1 . . . X", "S0 S1 . . .\n. . L0S . .\n. . . . G\n@ . . . .\nX . . . X"]
    ora, dev = make_pair(tiny, [e % 2 for e in range(2000)], 2000, seed=152)
    dev.vec.step(None); ora.step(None)  # a plain step first (thread-per-world kernel), then the loop (general kernel)
    _parts_rollout(ora, dev, 6, 40)
    dev.vec.step(None); ora.step(None)
    assert_same(dev, ora, dev.pull(), "tiny maps")
    ora, dev = make_pair([level_text(6)], None, 500, seed=153, obs_type="partial3x3", auto_reset=False)
    _parts_rollout(ora, dev, 4, 40)
    assert_same(dev, ora, dev.pull(), "partial observations")


def test_parts_loop_random_configurations():
    """Seeded sweep over batch sizes (ragged, down to one env), part counts (up to one ticket per part; fewer parts are formed when
    there are more parts than tickets), launch depths, levels, reward shapes and observation types."""
    rng = np.random.default_rng(20261019)
    for case in range(36):
        level = int(rng.integers(1, 7))
        n = int(rng.choice([1, 7, 31, 32, 33, 100, 257, 1000, 1999, 3000]))
        n_parts = int(rng.choice([1, 2, 3, 5, 8, 13, 32, 64, 200]))
        ahead = int(rng.integers(1, 5))
        kw = dict(seed=500 + case, reward_dim=int(rng.choice([1, 4])), auto_reset=bool(rng.integers(0, 2)),
                  obs_type=str(rng.choice(["layered", "layered", "partial3x3", "partial7x7", "perspective", "state"])))
        ora, dev = make_pair([level_text(level)], None, n, **kw)
        _parts_rollout(ora, dev, n_parts, 24, ahead)  # more parts than tickets: fewer parts are formed
        assert_same(dev, ora, dev.pull(), f"case {case}: level {level}, {n} envs, {n_parts} parts, depth {ahead}, {kw}")


def test_parts_loop_errors_and_abort():
    import lle_b200

    vec = lle_b200.VecWorld(level_text(6), 512, seed=1)
    n, A = 512, vec.n_agents
    act = torch.full((n, A), 4, dtype=torch.int8).pin_memory()
    rew = torch.empty((n, 1), dtype=torch.float32).pin_memory()
    done = torch.empty((n,), dtype=torch.uint8).pin_memory()
    with pytest.raises(ValueError):
        vec.parts_loop(4, torch.zeros((n, A), dtype=torch.int8), rew, done)  # not pinned
    with pytest.raises(ValueError):
        vec.parts_loop(0, act, rew, done)
    with pytest.raises(ValueError):
        vec.parts_loop(5000, act, rew, done)
    loop = vec.parts_loop(4, act, rew, done)
    with pytest.raises(ValueError):
        loop.wait(0)  # nothing fed
    with pytest.raises(IndexError):
        loop.feed(99)
    loop.feed(0)
    with pytest.raises(ValueError):
        loop.feed(0)  # the previous step of the part has not been waited for
    with pytest.raises(ValueError):
        loop.wait(0)  # fed, but no step launched
    with pytest.raises(ValueError):
        vec.step(None)  # the loop is open
    loop.launch()
    loop.wait(0)
    with pytest.raises(ValueError):
        loop.close()  # parts 1..3 of the launched step were never fed
    loop._open = True
    loop.abort()  # releases them (all STAY here), drains, closes
    vec.step(None)
    vec.synchronize()
    assert int(vec.err.sum()) == 0
    # a loop left open by an exception is aborted by the context manager; destroying a vec with an open loop does not hang
    with pytest.raises(RuntimeError):
        with vec.parts_loop(2, act, rew, done) as loop2:
            loop2.launch()
            raise RuntimeError("policy failed")
    vec.step(None)
    loop3 = vec.parts_loop(2, act, rew, done)
    loop3.launch()
    del loop3  # a dropped loop aborts itself: no launched step is left waiting for actions (it would block the whole device)
    vec.step(None)
    loop4 = vec.parts_loop(2, act, rew, done)
    loop4.launch()
    vec.destroy()  # lle_vec_destroy aborts an open loop first
    loop4._open = False
    other = lle_b200.VecWorld(level_text(3), 64, seed=2)  # the device is free again
    other.step(None)
    other.synchronize()


def test_veclle_run_host_policy():
    """VecLLE.run_host_policy: a host policy between steps; here it replays the oracle's sampled actions and checks what it is
    shown, part by part."""
    import lle_b200

    n, steps = 1500, 30
    ora = lo.OracleVec([level_text(6)], None, n, seed=77)
    rec = []
    for t in range(steps):
        ora.step(None)
        rec.append((np.array(ora.actions), np.array(ora.reward), np.array(ora.done)))
    env = lle_b200.VecLLE(level_text(6), n, seed=77)
    seen = {}

    def policy(sl, reward, done):
        t = seen.get(sl.start, 0)
        assert np.array_equal(reward, rec[t][1][sl]) and np.array_equal(done, rec[t][2][sl]), f"part at {sl.start}, step {t}"
        seen[sl.start] = t + 1
        return rec[t + 1][0][sl]

    rew, done = env.run_host_policy(policy, steps, n_parts=5, first_actions=rec[0][0])
    assert np.array_equal(rew.numpy(), rec[-1][1]) and np.array_equal(done.numpy(), rec[-1][2])
    env.world.synchronize()
    assert np.array_equal(env.obs.cpu().numpy(), np.asarray(ora.obs)) and int(env.err.sum()) == 0


def test_many_agents_and_small_maps():
    rows = [" .   .   . . . ."] + [f"S{k}  L{k}W  . . . X" for k in range(14)]
    run_pair(["\n".join(rows)], None, 64, 60, seed=2)
    run_pair(["S0 G X"], None, 64, 30, seed=3)          # 1x3 map: 18 floats per env, 32 envs per tile
    run_pair(["S0 . G\nS1 X X"], None, 70, 50, seed=4)  # 2x3 map


def test_large_map_chunked_tiles():
    """A 64x64 map with 8 agents and 16 sources: one observation (327,680 B) is streamed in chunks."""
    from _util import synthetic_map

    text = synthetic_map(64, 64, 8, 16, seed=5)
    run_pair([text], None, 64, 40, check_every=3, seed=6)


def _generated_maps():
    import json
    import os

    from _util import GOLDEN

    with open(os.path.join(GOLDEN, "generated_5x5.json")) as f:
        return json.load(f)["maps"]


def test_config3_generated_maps_heterogeneous_batch():
    """BASELINE configs[2]: 1,024 distinct generated 5x5 maps (2 agents, 2 lasers) in ONE batch — per-world static
    planes differ.  8 worlds per map against the oracle, bit-exact."""
    maps = _generated_maps()
    assert len(maps) == 1024 and len(set(maps)) == 1024
    per_map = 8
    moe = [m for m in range(1024) for _ in range(per_map)]
    run_pair(maps, moe, 1024 * per_map, 40, check_every=3, seed=21)
    # maps interleaved world by world (worst case for the tile cache: every sub-tile changes map)
    moe2 = [(e * 37) % 1024 for e in range(4096)]
    run_pair(maps, moe2, 4096, 25, check_every=4, seed=22)


def test_config3_full_size_properties():
    """1,024 maps x 1,024 worlds (1,048,576 worlds): size-independent checks."""
    import lle_b200

    maps = _generated_maps()
    n = 1024 * 1024
    moe = np.repeat(np.arange(1024, dtype=np.int32), 1024)
    vec = lle_b200.VecWorld(maps, n, map_of_env=moe, seed=5)
    A = vec.n_agents
    for t in range(30):
        vec.step(None)
    vec.synchronize()
    obs = vec.obs
    assert torch.all((obs == 0) | (obs == 1) | (obs == -1))
    assert torch.equal(obs[:, :A].sum(dim=(2, 3)), torch.ones(n, A, device=obs.device))
    pos = vec.state[:, : 2 * A].reshape(-1, A, 2).long()
    assert torch.equal(obs[:, :A].flatten(2).argmax(dim=2), pos[..., 0] * vec.width + pos[..., 1])
    assert torch.all(vec.err == 0)
    # worlds of one map share their static planes (walls, exits), worlds of different maps do not in general
    walls = obs[:, 2 * A].reshape(1024, 1024, -1)
    assert torch.equal(walls, walls[:, :1].expand_as(walls))
    assert len({tuple(w.tolist()) for w in walls[:, 0].cpu()}) > 500


def test_config3_full_size_windows_against_the_oracle():
    """BASELINE configs[2] at full size (1,024 maps x 1,024 worlds): an env's stream depends only on its global id, so
    windows of the 1,048,576-world batch are replayed by small oracle batches (env_id_base = window start) and compared
    bit-exactly, every output, for 40 steps."""
    import lle_b200

    maps = _generated_maps()
    n = 1024 * 1024
    moe = np.repeat(np.arange(1024, dtype=np.int32), 1024)
    vec = lle_b200.VecWorld(maps, n, map_of_env=moe, seed=6)
    rng = np.random.default_rng(3)
    starts = [0, n - 256] + [int(x) for x in rng.integers(0, n - 256, size=6)]  # windows straddle map boundaries too
    oracles = []
    for s0 in starts:
        ids = sorted(set(int(m) for m in moe[s0:s0 + 256]))
        local = [ids.index(int(m)) for m in moe[s0:s0 + 256]]
        oracles.append(lo.OracleVec([maps[i] for i in ids], local, 256, seed=6, env_id_base=s0))
    for t in range(40):
        vec.step(None)
        for ora in oracles:
            ora.step(None)
        if t % 5 == 4:
            vec.synchronize()
            for s0, ora in zip(starts, oracles):
                for name in ("obs", "state", "avail", "reward", "done", "events", "actions", "err"):
                    a = getattr(vec, name)[s0:s0 + 256].cpu().numpy()
                    assert np.array_equal(a, np.asarray(getattr(ora, name))), f"window at {s0}, step {t}: {name}"


def test_config5_large_batch_windows_against_the_oracle():
    """BASELINE configs[4] shape (64x64, 8 agents, 16 sources) at 16,384 worlds (5.4 GB of observations, 32 worlds per
    ticket): windows of the batch replayed by oracle batches with the same global env ids, bit-exact."""
    import lle_b200
    from _util import synthetic_map

    text = synthetic_map(64, 64, 8, 16, seed=5)
    n = 16384
    vec = lle_b200.VecWorld(text, n, seed=8)
    starts = [0, 5000, n - 24]
    oracles = [lo.OracleVec([text], None, 24, seed=8, env_id_base=s0) for s0 in starts]
    for t in range(24):
        vec.step(None)
        for ora in oracles:
            ora.step(None)
        if t % 6 == 5:
            vec.synchronize()
            for s0, ora in zip(starts, oracles):
                for name in ("obs", "state", "avail", "reward", "done", "events", "actions", "err"):
                    a = getattr(vec, name)[s0:s0 + 24].cpu().numpy()
                    assert np.array_equal(a, np.asarray(getattr(ora, name))), f"window at {s0}, step {t}: {name}"
    obs = vec.obs
    assert torch.equal(obs[:, :8].sum(dim=(2, 3)), torch.ones(n, 8, device=obs.device))  # one cell per agent plane, every world


def test_config4_mixed_levels_group():
    """BASELINE configs[3] at reduced size: the six levels mixed, one sub-batch per level, contiguous global env ids."""
    import lle_b200

    per_level = 512
    group = lle_b200.VecWorldGroup([(level_text(l), per_level) for l in range(1, 7)], seed=9)
    oracles = [lo.OracleVec([level_text(l)], None, per_level, seed=9, env_id_base=(l - 1) * per_level) for l in range(1, 7)]
    for t in range(60):
        group.step()
        for o in oracles:
            o.step(None)
    for part, o in zip(group.parts, oracles):
        d = Dev(part)
        assert_same(d, o, d.pull(), "mixed levels")


def test_set_state_fuzz(layouts):
    import ctypes as C

    rng = np.random.default_rng(42)
    L = lo.lib()
    for text in [level_text(6), level_text(5), layouts["eight-agent-interdependent-8"]]:
        n = 128
        ora, dev = make_pair([text], None, n, auto_reset=False, seed=9)
        H, W, A, G = ora.H, ora.W, ora.A, ora.G
        for rnd in range(4):
            ora.reset()
            dev.vec.reset()
            for _ in range(3 + rnd):
                ora.step(None)
                dev.vec.step(None)
            pos = np.stack([rng.integers(-1, H + 1, size=(n, A)), rng.integers(-1, W + 1, size=(n, A))], axis=-1).astype(np.int32)
            pos[: n // 2] = np.clip(pos[: n // 2], 0, [H - 1, W - 1])
            gems = rng.integers(0, 2, size=(n, G)).astype(np.uint8)
            alive = (rng.random((n, A)) < 0.85).astype(np.uint8)
            dev.vec.set_state(torch.from_numpy(pos), torch.from_numpy(gems), torch.from_numpy(alive))
            exp_err = np.zeros(n, np.uint8)
            for e in range(n):
                p = (C.c_long * (2 * A))(*[int(x) for x in pos[e].reshape(-1)])
                g = (C.c_uint8 * max(1, G))(*[int(x) for x in gems[e]])
                a = (C.c_uint8 * A)(*[int(x) for x in alive[e]])
                st = L.lleo_vec_set_state_env(ora._h, C.c_long(e), p, A, g, G, a)
                exp_err[e] = {0: 0, 105: 4, 104: 5}.get(st, 3 if st == 107 and b"same position" in L.lleo_last_error() else 6)
            L.lleo_vec_refresh(ora._h)
            raw = dev.pull()
            assert np.array_equal(dev.err, exp_err)
            for name in ("pos", "alive", "arrived", "slot", "collected"):
                assert np.array_equal(raw[name], np.asarray(getattr(ora, name))), name
            assert np.array_equal(raw["beam_on"][:, : ora.NB], ora.beam_on[:, : ora.NB])
            # `avail` included: after InvalidWorldState (world.rs:588-594) the reference returns before refreshing its
            # availability cache (:595); the device keeps the cache in the record and reproduces the stale values.
            for name in ("obs", "state", "done", "avail"):
                assert np.array_equal(getattr(dev, name), np.asarray(getattr(ora, name))), name


def test_rollout_mode_is_bit_identical_to_single_steps():
    """lle_vec_rollout(K): one launch, K steps; must equal K calls of step(None) and the oracle."""
    import lle_b200

    for level, n in ((6, 4096), (1, 1000), (5, 777)):
        ora, dev = make_pair([level_text(level)], None, n, seed=31)
        single = lle_b200.VecWorld(level_text(level), n, seed=31)
        done_steps = 0
        for k in (1, 7, 32, 5):
            dev.vec.rollout(k)
            for _ in range(k):
                ora.step(None)
                single.step(None)
            done_steps += k
            assert_same(dev, ora, dev.pull(), f"level {level} after {done_steps} rollout steps")
            assert dev.vec.step_count == single.step_count == done_steps
            assert torch.equal(dev.vec.obs, single.obs) and torch.equal(dev.vec.state, single.state)
    # heterogeneous maps + multi-objective in rollout mode
    maps = [level_text(2), level_text(3), level_text(4)]
    moe = [(e * 5 + e // 3) % 3 for e in range(640)]
    ora, dev = make_pair(maps, moe, 640, seed=12, reward_dim=4)
    for k in (16, 16, 3):
        dev.vec.rollout(k)
        for _ in range(k):
            ora.step(None)
        assert_same(dev, ora, dev.pull(), "heterogeneous rollout")


def test_determinism_and_shard_independence():
    """Same seed => same stream; an env's stream depends on its global id only, not on how envs are sharded."""
    import lle_b200

    text = level_text(6)
    full = lle_b200.VecWorld(text, 256, seed=3)
    lo_half = lle_b200.VecWorld(text, 128, seed=3, env_id_base=0)
    hi_half = lle_b200.VecWorld(text, 128, seed=3, env_id_base=128)
    for _ in range(60):
        for v in (full, lo_half, hi_half):
            v.step(None)
    full.synchronize()
    for name in ("obs", "state", "reward", "done", "events", "actions"):
        a = getattr(full, name).cpu()
        b = torch.cat([getattr(lo_half, name).cpu(), getattr(hi_half, name).cpu()])
        assert torch.equal(a, b), name


def test_golden_fixture_level6():
    """Committed fixture (tests/golden/rollout_lvl6.npz, generated by tests/golden/make_rollout_fixture.py)."""
    import os

    import lle_b200
    from _util import GOLDEN, rollout_digest

    fx = np.load(os.path.join(GOLDEN, "rollout_lvl6.npz"))
    vec = lle_b200.VecWorld(level_text(6), int(fx["n_envs"]), seed=int(fx["seed"]))
    dev = Dev(vec)
    for t in range(int(fx["steps"])):
        vec.step(None)
        dev.pull()
        assert rollout_digest(dev) == fx["digests"][t].tolist(), f"step {t}"
    assert np.array_equal(dev.obs, fx["final_obs"])
    assert np.array_equal(dev.state, fx["final_state"])


def test_full_size_properties():
    """Size-independent properties at BASELINE config 2 size (65,536 x lvl6), 200 steps."""
    import lle_b200

    vec = lle_b200.VecWorld(level_text(6), 65536, seed=99)
    m = vec.maps[0]
    A, C, H, W = vec.n_agents, vec.n_channels, vec.height, vec.width
    walls = torch.zeros(H, W)
    for i, j in m.walls:
        walls[i, j] = 1
    n_done = 0
    for t in range(200):
        vec.step(None)
        if t % 20 == 19:
            vec.synchronize()
            obs = vec.obs
            assert torch.all((obs == 0) | (obs == 1) | (obs == -1))
            assert torch.equal(obs[:, :A].sum(dim=(2, 3)), torch.ones(65536, A, device=obs.device))  # one cell per agent plane
            assert torch.equal(obs[:, 2 * A].cpu(), walls.expand(65536, H, W))                        # static wall plane
            st = vec.state
            pos = st[:, : 2 * A].reshape(-1, A, 2).long()
            onehot = obs[:, :A].flatten(2).argmax(dim=2)
            assert torch.equal(onehot, pos[..., 0] * W + pos[..., 1])                                 # state <-> obs
            gem_plane = obs[:, 2 * A + 2].sum(dim=(1, 2))
            assert torch.equal(gem_plane, (1 - st[:, 2 * A : 2 * A + vec.n_gems]).sum(dim=1))          # gems <-> obs
            assert torch.all(vec.err == 0)
            assert torch.all(vec.avail[:, :, 4] == 1)                                                 # STAY always available
            n_done += int(vec.done.sum())
    assert n_done > 0


@pytest.mark.gpu
def test_lle_facade_accessors(tmp_path):
    """The remaining members of `LLE` (python/lle/env/env.py:114-138, :218-254) on the N = 1 facade."""
    import lle_b200

    env = lle_b200.LLE.level(6)
    assert env.name == "LLE-lvl6" and (env.width, env.height) == (13, 12) and env.agent_state_size == 3
    assert env.world.n_agents == 4 and env.world.exit_pos == lle_b200.World.level(6).exit_pos
    assert (env.obs_type, env.state_type, env.walkable_lasers, env.randomize_lasers) == ("layered", "state", True, False)
    env.reset()
    obs, avail, extras = env.get_observation()
    assert obs.shape == (4, 12, 12, 13) and avail.shape == (4, 5) and extras.shape == (4, 0)
    assert env.compute_done() is False
    path = tmp_path / "tiny"
    path.write_text("S0 X")
    env = lle_b200.LLE.from_file(str(path), multi_objective=True)
    assert env.name == "LLE-tiny-MO" and env.reward_dim == 4  # builder.py:75
    env.reset()
    step = env.step([lle_b200.Action.EAST])
    assert step.done and env.compute_done() and lle_b200.LLE.from_str("S0 X").name == "LLE"
    assert lle_b200.level(1).name("mine").n_envs(4).build().name == "mine"
    # [P] python/tests/test_env.py test_env_name
    for lvl in (1, 6):
        assert lle_b200.level(lvl).build().name == f"LLE-lvl{lvl}" and lle_b200.level(lvl).multi_objective().build().name == f"LLE-lvl{lvl}-MO"
        assert lle_b200.LLE.level(lvl, multi_objective=True).name == f"LLE-lvl{lvl}-MO"
    assert lle_b200.from_str("S0 X").build().name == "LLE" and lle_b200.from_str("S0 X").multi_objective().build().name == "LLE-MO"
    assert lle_b200.from_file(str(path)).build().name == "LLE-tiny" and lle_b200.from_str("S0 L0E X").pbrs().build().name == "LLE-PBRS"
    # [P] python/tests/test_core.py test_width_height, test_state_default
    env = lle_b200.from_str("S0 X . .\n.  . . .\nG  . . .").build()
    assert (env.width, env.height) == (4, 3) and env.state_shape == (env.n_agents * 3 + 1,)
    env.reset()
    assert tuple(env.state.shape[1:]) == env.state_shape
    # python/tests/test_other.py:4-16
    env = lle_b200.from_str("X S0 G .").build()
    assert (env.width, env.height) == (4, 1)
    for lvl in range(1, 7):
        lle_b200.level(lvl).build()
    lle_b200.from_file(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "levels", "lvl1")).build()


def test_many_small_launches_in_flight():
    """A batch of one ticket: every step launch is a single CTA, so dozens of programmatically dependent launches are resident
    at once and the rotating scheduler slots are reused while their previous users still run (the slot generation word orders
    them).  Free-running for 4,000 steps, then every output against the oracle; and the same through rollout launches."""
    import lle_b200

    text = level_text(6)
    for n in (8, 64, 300):
        vec = lle_b200.VecWorld(text, n, seed=31)
        roll = lle_b200.VecWorld(text, n, seed=31)
        ora = lo.OracleVec([text], None, n, seed=31)
        for _ in range(4000):
            vec.step(None)
        for k in (1000, 1000, 2000):
            roll.rollout(k)
        for _ in range(4000):
            ora.step(None)
        d = Dev(vec)
        assert_same(d, ora, d.pull(), f"{n} envs, 4000 free-running steps")
        d = Dev(roll)
        assert_same(d, ora, d.pull(), f"{n} envs, rollout launches")


def test_dlpack_round_trip():
    """north_star: the batch is exposed as zero-copy DLPack / torch tensors.  torch.from_dlpack on the capsule of every buffer
    aliases the device buffer the kernel writes (same pointer; a step shows through the imported tensor)."""
    import lle_b200

    env = lle_b200.level(6).n_envs(256).seed(5).build()
    env.reset()
    imported = {name: torch.from_dlpack(env.dlpack(name)) for name in ("obs", "state", "available_actions", "reward", "done", "events", "actions")}
    for name, t in imported.items():
        src = getattr(env, name)
        assert t.data_ptr() == src.data_ptr() and t.shape == src.shape and t.dtype == src.dtype and t.device == src.device, name
    before = imported["state"].clone()
    for _ in range(5):
        env.step(None)
    torch.cuda.synchronize()
    assert not torch.equal(before, imported["state"])            # the imported tensor sees the kernel's writes
    assert torch.equal(imported["obs"], env.obs) and torch.equal(imported["state"], env.state)
    # numpy-style consumers on the host side of DLPack: a CPU copy round-trips through the protocol too
    host = torch.from_dlpack(env.state.cpu().__dlpack__())
    assert torch.equal(host, env.state.cpu())
    # the per-agent view (np.tile in the reference) is a stride-0 expand of the same memory
    per_agent = env.obs_per_agent
    assert per_agent.data_ptr() == env.obs.data_ptr() and per_agent.stride(1) == 0


def test_checkpoint_resume_is_bit_identical(layouts):
    """lle_vec_export_raw_state / lle_vec_import_raw_state: a fresh batch restored from a mid-rollout checkpoint continues exactly
    like the original — every output and the raw engine state, incl. stale beam bits after deaths (no auto-reset: the reference's
    WorldState could not carry them, world.rs:507-513), LaserSubgoal / PBRS flags, the reward counters, randomised laser colours."""
    import lle_b200

    cases = [
        dict(maps=[level_text(6)], n=300, kw=dict(seed=41)),
        dict(maps=[level_text(5)], n=200, kw=dict(seed=42, auto_reset=False, lle_semantics=False)),  # raw World: dead agents, stale bits
        dict(maps=[level_text(6)], n=200, kw=dict(seed=43, reward_dim=4, extras="laser_subgoal", pbrs=dict(gamma=0.9, reward_value=0.5))),
        dict(maps=[level_text(4)], n=150, kw=dict(seed=44, randomize_lasers=True)),
        dict(maps=[layouts["eight-agent-interdependent-8"]], n=100, kw=dict(seed=45, walkable_lasers=False)),
    ]
    for c in cases:
        a = lle_b200.VecWorld(c["maps"], c["n"], **c["kw"])
        for _ in range(37):
            a.step(None)
        ckpt = a.checkpoint()
        b = lle_b200.VecWorld(c["maps"], c["n"], **c["kw"])
        b.restore(ckpt)
        a.synchronize()
        for name in ("obs", "state", "avail"):  # re-exported from the imported records
            assert torch.equal(getattr(a, name), getattr(b, name)), f"{name} right after restore"
        for t in range(60):
            a.step(None)
            b.step(None)
            if t % 10 == 9:
                a.synchronize()
                for name in ("obs", "state", "avail", "reward", "done", "events", "actions", "err"):
                    assert torch.equal(getattr(a, name), getattr(b, name)), f"{name} differs {t + 1} steps after the restore"
                if a.extras is not None:
                    assert torch.equal(a.extras, b.extras)
                ra, rb = a.checkpoint(), b.checkpoint()
                for k in ra:
                    same = torch.equal(ra[k], rb[k]) if isinstance(ra[k], torch.Tensor) else ra[k] == rb[k]
                    assert same, f"raw '{k}' differs {t + 1} steps after the restore"


def test_episode_stats_and_info_against_the_oracle():
    """lle_vec_options.episode_stats: Step.info (env.py:174-188) and the episode return / length kept by the step kernel, against
    the same quantities accumulated on the host from the oracle's per-step outputs; general kernel and the tiny-map kernel."""
    import lle_b200

    for text, kw in ((level_text(6), dict(reward_dim=4)), (level_text(3), {}), ("S0 . G\nS1 X X", {}), ("S0 G . V S1\n.  . G . .\nX  . . G X", {})):
        n = 160
        vec = lle_b200.VecWorld(text, n, seed=51, episode_stats=True, **kw)
        ora = lo.OracleVec([text], None, n, seed=51, multi_objective=kw.get("reward_dim", 1) == 4)
        R, A = vec.reward_dim, vec.n_agents
        run_ret, run_len = np.zeros((n, R), np.float32), np.zeros(n, np.int32)
        last_ret, last_len = np.zeros((n, R), np.float32), np.zeros(n, np.int32)
        arrived = np.zeros((n, A), bool)
        for t in range(150):
            vec.step(None)
            ora.step(None)
            rew, done, ev = np.asarray(ora.reward), np.asarray(ora.done).astype(bool), np.asarray(ora.events)
            run_ret += rew
            run_len += 1
            arrived |= (ev & 3) == 1  # AgentExit in the first pass (later passes only emit deaths)
            vec.synchronize()
            info = vec.info
            assert np.array_equal(info["exit_rate"].cpu().numpy(), arrived.sum(1).astype(np.float32) / A), f"exit_rate, step {t}"
            for i in range(A):
                assert np.array_equal(info[f"has-arrived-{i}"].cpu().numpy(), arrived[:, i]), f"has-arrived-{i}, step {t}"
            last_ret[done], last_len[done] = run_ret[done], run_len[done]
            run_ret[done], run_len[done], arrived[done] = 0, 0, False
            assert np.array_equal(vec.ep_return.cpu().numpy(), run_ret) and np.array_equal(vec.ep_length.cpu().numpy(), run_len), f"running, step {t}"
            assert np.array_equal(vec.last_return.cpu().numpy(), last_ret) and np.array_equal(vec.last_length.cpu().numpy(), last_len), f"last, step {t}"
        assert last_len.max() > 0
        # gems_collected: the N = 1 facade's info (raw engine state) is the reference for the batch
    text = "S0 G . V S1\n.  . G . .\nX  . . G X"
    env = lle_b200.LLE(text)
    vec = lle_b200.VecWorld(text, 1, episode_stats=True, auto_reset=False)
    env.reset()
    rng = np.random.default_rng(3)
    seen = set()
    for _ in range(120):
        mask = env.available_actions()
        acts = [int(rng.choice(np.flatnonzero(mask[a]))) for a in range(2)]
        step = env.step(acts)
        vec.step(torch.tensor([acts], dtype=torch.int8))
        vec.synchronize()
        assert int(vec.info["gems_collected"][0]) == step.info["gems_collected"]
        assert float(vec.info["exit_rate"][0]) == step.info["exit_rate"]
        seen.add(step.info["gems_collected"])
        if step.done:
            env.reset()
            vec.reset()
    assert len(seen) >= 2
