"""Known-answer tests of the layered observation, the state vector, rewards and done flags,
transcribed from the reference:

  [O]  python/tests/test_observations.py
  [E]  python/tests/test_env.py
  [C]  python/tests/test_core.py
  [S]  python/tests/test_reward_strategy.py

Run against the oracle (CPU) and, with ``-m gpu``, against the CUDA path (``api.LLE`` is then the
N=1 view of the batched environment).
"""
import numpy as np
import pytest


def channels(n_agents):
    """python/lle/observations.py:205-211."""
    a0, laser0 = 0, n_agents
    wall = laser0 + n_agents
    return dict(A0=a0, LASER_0=laser0, WALL=wall, VOID=wall + 1, GEM=wall + 2, EXIT=wall + 3, C=wall + 4)


# ----------------------------------------------------------------------------- layered observation
def test_observe_layered_deactivated_laser(api):  # [O]
    world = api.World(
        """
@ @ L0S @  @
@ .  .  .  @
@ X  .  S0 @
@ X  .  S1 @
@ @  @  @  @
"""
    )
    ch = channels(2)
    world.reset()
    layers = world.observe_layered()
    assert layers.shape == (2, ch["C"], 5, 5) and layers.dtype == np.float32
    l0, l1 = ch["LASER_0"], ch["LASER_0"] + 1
    assert np.all(layers[:, l0, 0, 2] == -1)
    assert np.all(layers[:, l0, 1:4, 2] == 1)
    assert np.all(layers[:, l1] == 0)
    world.step([api.Action.WEST, api.Action.STAY])
    layers = world.observe_layered()
    assert np.all(layers[:, l0, 0, 2] == -1)
    assert np.all(layers[:, l0, 1, 2] == 1)
    assert np.all(layers[:, l0, 2:4, 2] == 0)
    assert np.all(layers[:, l1] == 0)
    # agent layers (observations.py:264-265)
    assert layers[0, ch["A0"], 2, 2] == 1 and layers[0, ch["A0"]].sum() == 1
    assert layers[0, ch["A0"] + 1, 3, 3] == 1 and layers[0, ch["A0"] + 1].sum() == 1


def test_observe_layered_gems_walls(api):  # [O]
    world = api.World(
        """
@ @ L0S @  @
@ .  .  .  @
@ X  G  S0 @
@ .  .  .  @
@ @  @  @  @
"""
    )
    ch = channels(1)
    world.reset()
    layers = world.observe_layered()
    for i, j in world.wall_pos:
        assert np.all(layers[:, ch["WALL"], i, j] == 1)
    assert layers[0, ch["WALL"]].sum() == len(world.wall_pos)
    for gem in world.gems:
        assert np.all(layers[:, ch["GEM"], gem.pos[0], gem.pos[1]] == 1)
    for i, j in world.exit_pos:
        assert np.all(layers[:, ch["EXIT"], i, j] == 1)
    for laser in world.lasers:
        if laser.is_on:
            assert np.all(layers[:, ch["LASER_0"] + laser.agent_id, laser.pos[0], laser.pos[1]] == 1)
    for source in world.laser_sources:
        assert np.all(layers[:, ch["LASER_0"] + source.agent_id, source.pos[0], source.pos[1]] == -1)
    assert np.all(layers[:, ch["VOID"]] == 0)


def test_observe_layered_void(api):  # [O]
    world = api.World(
        """
    V . . S0
    . . . .
    V V G X"""
    )
    ch = channels(1)
    world.reset()
    layers = world.observe_layered()
    expected = np.zeros((3, 4), dtype=np.float32)
    for i, j in [(0, 0), (2, 0), (2, 1)]:
        expected[i, j] = 1.0
    assert np.array_equal(layers[0, ch["VOID"]], expected)


def test_observe_gem_layer_follows_collection(api):  # observations.py:260-263
    world = api.World("S0 G X")
    ch = channels(1)
    world.reset()
    assert world.observe_layered()[0, ch["GEM"], 0, 1] == 1
    world.step([api.Action.EAST])
    assert world.observe_layered()[0, ch["GEM"]].sum() == 0


def test_layered_laser_colour_above_n_agents(api):  # [O] test_layered_observation_laser_source_agent_id_above_n_agents
    world = api.World("S0 L1E X")
    ch = channels(1)
    data = world.observe_layered()
    laser_1_layer = data[0, ch["LASER_0"] + 1]  # aliases the WALL channel
    assert laser_1_layer[0, 1] == -1
    assert laser_1_layer[0, 2] == 1


def test_layered_laser_colour_out_of_channels(api):  # numpy IndexError in observations.py:235
    world = api.World("S0 L9E X")
    with pytest.raises(IndexError):
        world.observe_layered()


def test_all_shapes(api):  # [O] test_all_shapes (layered / flattened / state)
    for level in range(1, 7):
        world = api.World.level(level)
        obs = world.observe_layered()
        a = world.n_agents
        assert obs.shape == (a, 2 * a + 4, 12, 13)
        assert obs.reshape(a, -1).shape == (a, (2 * a + 4) * 12 * 13)  # FlattenedLayered (observations.py:291-293)
        assert world.state_array().shape == (3 * a + world.n_gems,)
        for k in range(1, a):
            assert np.array_equal(obs[0], obs[k])  # np.tile (observations.py:266)


def test_level6_initial_observation(api):  # Appendix D of SURVEY.md + resources/levels/lvl6
    world = api.World.level(6)
    world.reset()
    ch = channels(4)
    obs = world.observe_layered()[0]
    for a in range(4):
        assert obs[ch["A0"] + a, 0, 4 + a] == 1 and obs[ch["A0"] + a].sum() == 1
    # L2S at (0,1): beam of 2 cells stopped by the wall at (3,1)
    assert obs[ch["LASER_0"] + 2, 0, 1] == -1 and obs[ch["LASER_0"] + 2, 1:3, 1].tolist() == [1, 1]
    assert obs[ch["LASER_0"] + 2].sum() == 1
    # L0E at (4,0): 6 cells up to the wall at (4,7)
    assert obs[ch["LASER_0"], 4, 0] == -1 and obs[ch["LASER_0"], 4, 1:7].tolist() == [1] * 6
    # L1W at (6,12): 12 cells
    assert obs[ch["LASER_0"] + 1, 6, 12] == -1 and obs[ch["LASER_0"] + 1, 6, 0:12].tolist() == [1] * 12
    assert obs[ch["GEM"]].sum() == 4 and obs[ch["EXIT"]].sum() == 4 and obs[ch["VOID"]].sum() == 0
    assert obs[ch["WALL"]].sum() == len(world.wall_pos) == 18  # 15 walls + 3 sources (parser_v1.rs:22-25)


def test_observation_gem_collected_state(api):  # [O] test_observation_gem_collected
    world = api.World("S0 X . .\n.  . . .\nG  . . .")
    world.reset()
    world.step([api.Action.SOUTH])
    assert world.state_array()[2] == 0.0
    world.step([api.Action.SOUTH])
    assert world.state_array()[2] == 1.0
    world.step([api.Action.NORTH])
    assert world.state_array()[2] == 1.0
    assert world.state_array().tolist() == [1.0, 0.0, 1.0, 1.0]


# ----------------------------------------------------------------------------- rewards / done (LLE)
def test_void_reward(api):  # [E]
    env = api.LLE("S0 V X")
    env.reset()
    step = env.step([api.Action.EAST])
    assert step.reward.tolist() == [-1.0]
    assert step.done


def test_collect_reward(api):  # [E]
    env = api.LLE("S0 X . .\n.  . . .\nG  . . .")
    env.reset()
    env.step([api.Action.SOUTH])
    assert env.step([api.Action.SOUTH]).reward.tolist() == [1.0]


def test_time_reward(api):  # [E]
    env = api.LLE(". .  . X\n. S0 . .\n. .  . .")
    env.reset()
    for action in api.Action:
        assert env.step([action]).reward.tolist() == [0.0]


def test_finish_reward(api):  # [E]
    env = api.LLE(
        """@ @ @  @ @ @
@ . .  . . @
@ . S0 . . @
@ . .  X . @
@ @ @  @ @ @"""
    )
    env.reset()
    env.step([api.Action.EAST])
    step = env.step([api.Action.SOUTH])
    assert step.reward.tolist() == [2.0] and step.done


def test_arrive_reward_only_once(api):  # [E] 7-step trace
    A = api.Action
    env = api.LLE("S0 . G\nS1 X X")
    trace = [
        ([A.EAST, A.STAY], 0),
        ([A.STAY, A.EAST], 1),
        ([A.STAY, A.STAY], 0),
        ([A.STAY, A.STAY], 0),
        ([A.EAST, A.STAY], 1),
        ([A.STAY, A.STAY], 0),
        ([A.SOUTH, A.STAY], 2),
    ]
    env.reset()
    for k, (action, reward) in enumerate(trace):
        step = env.step(action)
        assert step.reward.tolist() == [float(reward)]
        assert step.done == (k == len(trace) - 1)


def test_reward_after_reset(api):  # [E]
    A = api.Action
    env = api.LLE("S0 X . .\n.  . . .\nG  . . .")
    for _ in range(10):
        env.reset()
        env.step([A.SOUTH])
        assert env.step([A.SOUTH]).reward.tolist() == [1.0]
        assert not env.done
        assert env.step([A.NORTH]).reward.tolist() == [0.0]
        assert env.step([A.NORTH]).reward.tolist() == [0.0]
        step = env.step([A.EAST])
        assert env.done and step.reward.tolist() == [2.0]


def test_reward_after_set_state(api):  # [E]
    env = api.LLE("S0 . G\nS1 X X")
    env.reset()
    env.set_state(api.WorldState([(0, 1), (1, 1)], [False]))
    assert env.step([api.Action.EAST, api.Action.STAY]).reward.tolist() == [1.0]


def test_reward_set_state_all_arrived(api):  # [E]
    env = api.LLE("S0 . G\nS1 X X")
    env.reset()
    env.set_state(api.WorldState([(0, 2), (1, 1)], [True]))
    assert env.step([api.Action.SOUTH, api.Action.STAY]).reward.tolist() == [2.0]


def test_reward(api):  # [E]
    A = api.Action
    env = api.LLE("S0 G .\n.  . X")
    env.reset()
    assert env.step([A.EAST]).reward.tolist() == [1.0]
    assert env.step([A.EAST]).reward.tolist() == [0.0]
    assert env.step([A.SOUTH]).reward.tolist() == [2.0]


def test_reward_death(api):  # [E]
    env = api.LLE("S0 L0S X\nS1  .  X")
    env.reset()
    step = env.step([api.Action.STAY, api.Action.EAST])
    assert step.reward.tolist() == [-1.0]
    assert step.done


def test_reward_collect_and_death(api):  # [E]
    env = api.LLE("S0 L0S X\nS1  G  X")
    env.reset()
    step = env.step([api.Action.STAY, api.Action.EAST])
    assert step.reward.item() == -1.0
    assert step.done


def test_step_in_done_env_raises(api):  # env.py:166-167
    env = api.LLE("S0 V X")
    env.reset()
    env.step([api.Action.EAST])
    with pytest.raises(ValueError):
        env.step([api.Action.STAY])


def test_multi_objective_rewards(api):  # [E]
    A = api.Action
    env = api.LLE("S0 G .\n.  . X", multi_objective=True)
    env.reset()
    assert env.step([A.EAST]).reward.tolist() == [1.0, 0.0, 0.0, 0.0]
    assert env.step([A.EAST]).reward.tolist() == [0.0, 0.0, 0.0, 0.0]
    step = env.step([A.SOUTH])
    assert step.done
    assert step.reward.tolist() == [0.0, 1.0, 0.0, 1.0]


def test_multi_objective_death(api):  # [E]
    env = api.LLE("S0 L0S X\nS1  G  X", multi_objective=True)
    env.reset()
    assert env.step([api.Action.STAY, api.Action.EAST]).reward.tolist() == [0.0, 0.0, -1.0, 0.0]


def test_env_available_actions(api):  # [C] test_available_actions
    A = api.Action
    env = api.LLE(
        """
@ @ L0S @  @
@ .  .  .  @
@ X  .  S0 @
@ X  .  S1 @
@ @  @  @  @
"""
    )
    obs, state = env.reset()
    av = env.available_actions()
    assert av.shape == (2, 5) and av.dtype == bool
    assert av[0].tolist() == [True, False, False, True, True]  # N, S, E, W, STAY
    assert av[1].tolist() == [False, False, False, True, True]


WALK = """
@ @ L{c}S @  @
@ .  .  .  @
@ X  {m0} {r0} @
@ X  .  {r1} @
@ @  @  @  @
"""


def test_walkable_lasers(api):  # python/tests/test_walkable_lasers.py (all five cases) ; env.py:153-163
    W = api.Action.WEST
    # enabled (default): agents may walk into any laser
    env = api.LLE(WALK.format(c=0, m0=".", r0="S0", r1="S1"))
    env.reset()
    av = env.available_actions()
    assert av[0, W] and av[1, W]
    # disabled, laser active: the other colour may not enter
    env = api.LLE(WALK.format(c=0, m0=".", r0="S0", r1="S1"), walkable_lasers=False)
    env.reset()
    av = env.available_actions()
    assert av[0, W] and not av[1, W]
    # disabled, beam blocked by its owner standing in it
    env = api.LLE(WALK.format(c=0, m0="S0", r0=".", r1="S1"), walkable_lasers=False)
    env.reset()
    assert env.available_actions()[1, W]
    # switched colours
    env = api.LLE(WALK.format(c=1, m0=".", r0="S0", r1="S1"), walkable_lasers=False)
    env.reset()
    av = env.available_actions()
    assert not av[0, W] and av[1, W]
    env = api.LLE(WALK.format(c=1, m0="S1", r0=".", r1="S0"), walkable_lasers=False)
    env.reset()
    assert env.available_actions()[0, W]


def test_force_end_state_env(api):  # [C] test_force_end_state
    env = api.LLE("S0 . G\nX  . .")
    env.reset()
    env.set_state(api.WorldState([(1, 0)], [True]))
    assert env.done


def test_move_end_game(api):  # [C] test_move_end_game
    A = api.Action
    env = api.LLE(
        """
    S0 X .
    .  . .
    .  . ."""
    )
    env.reset()
    env.step([A.SOUTH])
    assert not env.done
    env.step([A.SOUTH])
    assert not env.done
    env.step([A.EAST])
    assert not env.done
    env.step([A.NORTH])
    assert not env.done
    step = env.step([A.NORTH])
    assert step.done and env.done


# ----------------------------------------------------------------------------- LaserSubgoal extras / PBRS
PBRS_MAP = "S0 .  .\n.  . L0W\nX  .  ."
TWO_LASERS = "S0  S1 X  X\n.   .  . L0W\n.   .  . L1W"


def test_pbrs_reset_between_two_episodes(api):  # [E]
    A = api.Action
    actions = [[A.SOUTH], [A.EAST], [A.SOUTH], [A.WEST]]
    expected = [1.0, 0.0, 0.0, 2.0]  # SHAPED_REWARD, 0, 0, REWARD_EXIT + REWARD_DONE
    env = api.LLE(PBRS_MAP, pbrs=dict(reward_value=1.0, gamma=1.0))
    for _ in range(5):
        env.reset()
        for action, reward in zip(actions, expected):
            assert env.step(action).reward.tolist() == [reward]


def test_pbrs_not_all_lasers(api):  # [E] .pbrs(lasers_to_reward=[(1, 2)])
    text = "S0 .  .\n.  . L0W\n.  . L0W\nX  .  ."
    world = api.World(text)
    idx = [tuple(s.pos) for s in world.laser_sources].index((1, 2))
    env = api.LLE(text, pbrs=dict(lasers_to_reward=[idx]))
    env.reset()
    assert env.extras().shape == (1, 1)


def _extras_one_agent(env):
    env.reset()
    extras = env.extras()
    assert extras.shape == (1, 1) and extras.dtype == np.float32
    assert extras[0][0] == 0.0


@pytest.mark.parametrize("kw", [dict(extras="laser_subgoal"), dict(pbrs=dict(with_extras=True))], ids=["extras", "pbrs"])
def test_subgoal_extras_one_laser(api, kw):  # [O] test_subgoal_extras_one_laser / test_pbrs_subgoals_extras_one_laser
    env = api.LLE("S0  X\n.  L0W", **kw)
    _extras_one_agent(env)
    env.reset()
    _extras_one_agent(env)


def _extras_two_agents(env, A):  # [O] _perform_tests_two_agents
    env.reset()
    extras = env.extras()
    assert extras.shape == (2, 2) and np.all(extras == 0.0)
    step = env.step([A.SOUTH, A.STAY])
    assert step.extras.shape == (2, 2)
    assert step.extras[0].sum() == 1.0 and step.extras[1].sum() == 0.0
    step = env.step([A.NORTH, A.STAY])
    assert step.extras[0].sum() == 1.0 and step.extras[1].sum() == 0.0
    # even when an agent dies, the subgoal is reached
    step = env.step([A.STAY, A.SOUTH])
    assert step.done
    assert step.extras[0].sum() == 1.0 and step.extras[1].sum() == 1.0


@pytest.mark.parametrize("kw", [dict(extras="laser_subgoal"), dict(pbrs=dict(with_extras=True))], ids=["extras", "pbrs"])
def test_subgoal_extras_two_lasers_two_agents(api, kw):  # [O]
    env = api.LLE(TWO_LASERS, **kw)
    _extras_two_agents(env, api.Action)
    env.reset()
    _extras_two_agents(env, api.Action)


def test_pbrs_single_objective(api):  # [S] test_pbrs_single_objective / test_pbrs_with_lle
    A = api.Action
    env = api.LLE(PBRS_MAP, pbrs=dict(gamma=0.99, reward_value=0.5))
    env.reset()
    step = env.step([A.EAST])  # no base reward; the potential is unchanged
    assert step.reward.dtype == np.float32 and step.reward.shape == (1,)
    assert step.reward[0] == np.float32(0.5 * 0.99 - 0.5)
    step = env.step([A.SOUTH])  # into the laser: the potential drops to 0
    assert step.reward[0] == np.float32(0.5 * 0.99 - 0)
    step = env.step([A.SOUTH])
    assert step.reward[0] == 0.0


def test_pbrs_multi_objective(api):  # [S]
    A = api.Action
    env = api.LLE(PBRS_MAP, multi_objective=True, pbrs=dict(gamma=0.99, reward_value=0.5))
    env.reset()
    step = env.step([A.EAST])
    assert step.reward.shape == (5,)
    assert np.allclose(step.reward, np.array([0.0] * 4 + [0.99 * 0.5 - 0.5], dtype=np.float32))
    step = env.step([A.SOUTH])
    assert np.allclose(step.reward, np.array([0.0] * 4 + [0.5 * 0.99], dtype=np.float32))
    step = env.step([A.SOUTH])
    assert step.reward[-1] == 0.0


def test_pbrs_builder_order():  # [S] test_pbrs_raises_value_error (host-side builder logic, no device needed)
    import lle_b200

    lle_b200.from_str(PBRS_MAP).multi_objective().pbrs()
    with pytest.raises(ValueError):
        lle_b200.from_str(PBRS_MAP).pbrs().multi_objective()
    with pytest.raises(ValueError):
        lle_b200.from_str(PBRS_MAP).pbrs(lasers_to_reward=[(0, 0)])
    with pytest.raises(ValueError):
        lle_b200.from_str(PBRS_MAP).add_extras("nope")


# ----------------------------------------------------------------------------- other observation types
# The reference builds a generator on a World (`PartialGenerator(world, 3)`); here the same generator is selected with
# `LLE(map, obs_type=...)` (Builder.obs_type, builder.py:42-49) and `env.reset()` stands for `world.reset()`.
def test_observe_flattened(api):  # [O]
    text = "@ @ L0S @  @\n@ .  .  .  @\n@ X  G  S0 @\n@ .  .  .  @\n@ @  @  @  @"
    env = api.LLE(text, obs_type="flattened")
    obs, _ = env.reset()
    assert obs.shape == (1, (1 * 2 + 4) * 5 * 5)
    layered, _ = api.LLE(text).reset()
    assert np.array_equal(obs, layered.reshape(1, -1))


def test_world_initial_observation_normalized_state(api):  # [O] test_world_initial_observation
    obs, _ = api.LLE("S0 X .\n.  . .\n.  . .", obs_type="normalized-state").reset()
    assert np.array_equal(np.array([[0.0, 0.0, 1.0]]), obs)
    obs, _ = api.LLE("S0 X  .\n.  .  S1\n.  .  X", obs_type="normalized-state").reset()
    assert obs.dtype == np.float32
    assert np.allclose(np.tile(np.array([0.0, 0.0, 1 / 3, 2 / 3, 1.0, 1.0]), (2, 1)), obs)
    # float32 state / int64 dimensions -> float64 quotient stored back into float32 (observations.py:156-157)
    assert np.array_equal(obs[0], np.array([0.0, 0.0, 1.0 / 3, 2.0 / 3, 1.0, 1.0]).astype(np.float32))
    obs, _ = api.LLE("S0 X  .  .\n.  .  S1  .\n.  X  .  .", obs_type="normalized-state").reset()
    assert np.allclose(np.tile(np.array([0.0, 0.0, 1 / 3, 1 / 2, 1.0, 1.0]), (2, 1)), obs)
    obs, _ = api.LLE("S0 X  .  G\n.  .  S1  .\n.  X  .  .", obs_type="normalized-state").reset()
    assert np.allclose(np.tile(np.array([0.0, 0.0, 1 / 3, 1 / 2, 0.0, 1.0, 1.0]), (2, 1)), obs)


def test_state_observation_type(api):  # [O] test_observation_gem_collected through the env
    env = api.LLE("S0 X . .\n.  . . .\nG  . . .", obs_type="state")
    obs, state = env.reset()
    assert obs.shape == (1, 4) and np.array_equal(obs[0], state)
    env.step([api.Action.SOUTH])
    obs = env.step([api.Action.SOUTH]).obs
    assert np.all(obs[:, 2] == 1.0)


def test_partial_3x3(api):  # [O]
    env = api.LLE("S0 X  @\nG  S1 @\n.  .  X", obs_type="partial3x3")
    obs, _ = env.reset()
    WALL, LASER_0 = 2, 3
    GEM, EXIT = LASER_0 + 2, LASER_0 + 3
    assert obs.shape == (2, 7, 3, 3) and obs.dtype == np.float32
    obs0, obs1 = obs
    assert obs0[0, 1, 1] == 1 and obs0[1, 2, 2] == 1
    assert obs1[0, 0, 0] == 1 and obs1[1, 1, 1] == 1
    assert obs0[GEM, 2, 1] == 1 and obs1[GEM, 1, 0] == 1
    assert obs1[EXIT, 2, 2] == 1
    assert np.all(obs0[WALL] == 0)
    assert obs1[WALL, 1, 2] == 1 and obs1[WALL, 0, 2] == 1


def test_partial_7x7(api):  # [O]
    env = api.LLE("S0 S1 S2 S3 X X X X", obs_type="partial7x7")
    observations, _ = env.reset()
    n_agents, center = 4, 3
    WALL, LASER_0 = n_agents, n_agents + 1
    GEM, EXIT = LASER_0 + n_agents, LASER_0 + n_agents + 1
    assert observations.shape == (4, 2 * n_agents + 3, 7, 7)
    observations = observations.copy()
    for agent_num, obs in enumerate(observations):
        for other in range(n_agents):
            i, j = center, center - agent_num + other  # agents are side by side
            assert obs[other, i, j] == 1
            obs[other, i, j] = 0
            assert np.all(obs[0] == 0)
    assert np.all(observations[0, EXIT] == 0)
    assert observations[1, EXIT, center, center + 3] == 1
    assert np.all(observations[2, EXIT, center, center + 2:] == 1)
    assert np.all(observations[3, EXIT, center, center + 1:] == 1)
    assert np.all(observations[:, WALL] == 0)
    assert np.all(observations[:, GEM] == 0)
    assert np.all(observations[:, LASER_0:LASER_0 + n_agents] == 0)


def test_partial_3x3_lasers(api):  # [O]
    env = api.LLE(".   L0S S1\nS0   .   .\nL1E  X   X", obs_type="partial3x3")
    obs, _ = env.reset()
    LASER_0 = 3
    obs0, obs1 = obs
    assert obs0[LASER_0, 0, 2] == -1
    assert obs0[LASER_0, 1, 2] == 1
    assert obs0[LASER_0, 2, 2] == 1
    assert obs0[LASER_0 + 1, 2, 1] == -1
    assert obs0[LASER_0 + 1, 2, 2] == 1


def test_partial_5x5_shape_and_off_map_cells(api):  # observations.py:301, :325-329
    env = api.LLE("S0 X", obs_type="partial5x5")
    obs, _ = env.reset()
    assert obs.shape == (1, 5, 5, 5)
    assert obs[0, 0, 2, 2] == 1 and obs[0, 4, 2, 3] == 1  # the agent in the centre, the exit to its right
    assert obs.sum() == 2  # everything outside the 1x2 map is 0


def test_padded_layered(api):  # [O]
    base, _ = api.LLE("S0 X").reset()
    for k in (1, 2, 3):
        obs, _ = api.LLE("S0 X", obs_type=f"layered-padded-{k}").reset()
        assert obs.shape[1] == base.shape[1] + 2 * k and obs.shape[2:] == base.shape[2:]
        assert obs.shape[0] == 1 + k  # np.tile over the padded agent count (observations.py:203, :266)
        # the padded agent and laser layers are empty; the others keep their order
        assert np.array_equal(obs[0, 0], base[0, 0]) and np.all(obs[:, 1:1 + k] == 0)
        assert np.array_equal(obs[0, 1 + k], base[0, 1]) and np.all(obs[:, 2 + k:2 + 2 * k] == 0)
        assert np.array_equal(obs[0, 2 + 2 * k:], base[0, 2:])
    obs, _ = api.LLE("S0 X", obs_type="layered-padded", padding_size=5).reset()
    assert obs.shape == (6, 16, 1, 2)


def test_padded_layered_takes_foreign_colours(api):  # colour >= n_agents lands in a padded laser layer instead of WALL
    obs, _ = api.LLE("S0 L1E X", obs_type="layered-padded-1").reset()
    LASER_0 = 2
    assert obs[0, LASER_0 + 1, 0, 1] == -1 and obs[0, LASER_0 + 1, 0, 2] == 1


def test_perspective(api):  # [O]
    env = api.LLE("S0  S1 S2 X\nL0E .  X  .\n .  .  X L1W", obs_type="perspective")
    obs, _ = env.reset()
    A0, L0 = 0, 3
    assert obs.shape == (3, 10, 3, 4)
    obs0, obs1, obs2 = obs
    assert obs0[A0, 0, 0] == 1 and obs1[A0, 0, 1] == 1 and obs2[A0, 0, 2] == 1
    assert obs0[L0, 1, 0] == -1
    assert np.all(obs0[L0, 1, 1:] == 1)
    assert obs1[L0, 2, 3] == -1
    assert np.all(obs1[L0, 2, :3] == 1)


def test_perspective2(api):  # [O]
    text = "S0  S1 S2\n .   .  .\nL0E  X  .\nL1E  X  .\nL2E  X  ."
    base, persp = api.LLE(text), api.LLE(text, obs_type="perspective")
    layered_obs, _ = base.reset()
    perspective_obs, _ = persp.reset()
    A0, L0 = 0, 3
    positions = [(0, 0), (0, 1), (0, 2)]
    for actions in (None, [api.Action.SOUTH] * 3):
        if actions is not None:
            layered_obs = base.step(actions).obs
            perspective_obs = persp.step(actions).obs
            positions = [(i + 1, j) for i, j in positions]
        assert perspective_obs.shape == (3, 10, 5, 3)
        for observer, position in enumerate(positions):
            expected = np.copy(layered_obs[observer])
            expected[[A0, A0 + observer]] = expected[[A0 + observer, A0]]
            expected[[L0, L0 + observer]] = expected[[L0 + observer, L0]]
            np.testing.assert_array_equal(perspective_obs[observer], expected)
            assert perspective_obs[observer, A0, position[0], position[1]] == 1.0


def test_all_observation_shapes(api):  # [O] test_all_shapes for the generators on the accelerated path
    expect = {"layered": lambda A, G: (2 * A + 4, 12, 13), "flattened": lambda A, G: ((2 * A + 4) * 12 * 13,),
              "partial3x3": lambda A, G: (2 * A + 3, 3, 3), "partial5x5": lambda A, G: (2 * A + 3, 5, 5),
              "partial7x7": lambda A, G: (2 * A + 3, 7, 7), "state": lambda A, G: (3 * A + G,),
              "normalized-state": lambda A, G: (3 * A + G,), "perspective": lambda A, G: (2 * A + 4, 12, 13),
              "layered-padded-1": lambda A, G: (2 * A + 6, 12, 13), "layered-padded-2": lambda A, G: (2 * A + 8, 12, 13),
              "layered-padded-3": lambda A, G: (2 * A + 10, 12, 13)}
    for level in (1, 3, 6):
        for name, shape in expect.items():
            env = api.LLE.level(level, obs_type=name)
            obs, _ = env.reset()
            A, G = env.n_agents, env.n_gems
            pad = int(name[-1]) if name.startswith("layered-padded-") else 0
            assert obs.shape == (A + pad, *shape(A, G)), (level, name, obs.shape)


def test_unknown_observation_type(api):
    with pytest.raises(ValueError):
        api.LLE("S0 X", obs_type="nope")


def test_randomized_lasers(api):  # [E] python/tests/test_env.py:381 test_randomized_lasers
    env = api.LLE("S0 S1 L0S\n.   . L1W\n.   . L0W\nX   X  .", randomize_lasers=True)
    n_sources = len(env.laser_sources)
    encountered = [[False] * env.n_agents for _ in range(n_sources)]
    for _ in range(200):  # the reference allows 1,000 tries; 2^-200 is already impossible enough
        env.reset()
        for k, source in enumerate(env.laser_sources):
            encountered[k][source.agent_id] = True
        if all(all(e) for e in encountered):
            break
    assert all(all(e) for e in encountered), "the two colours were never encountered for some lasers"
    # the laser tiles follow their source (world.lasers reads the live colour)
    by_id = {s.laser_id: s.agent_id for s in env.laser_sources}
    assert all(l.agent_id == by_id[l.laser_id] for l in env.lasers)


def test_deep_copy_env(api):  # [P] python/tests/test_core.py test_deep_copy
    from copy import deepcopy

    env = api.LLE("S0 X")
    other = deepcopy(env)
    assert env is not other
    env.reset()
    other.reset()
    assert env.step([api.Action.EAST.value]).done
    assert not other.step([api.Action.STAY.value]).done   # the copy is an environment of its own
    # a copy taken in the middle of an episode continues from there
    env = api.LLE("S0 G . X")
    env.reset()
    env.step([api.Action.EAST.value])
    other = deepcopy(env)
    assert np.array_equal(other.get_state(), env.get_state())
    assert other.step([api.Action.EAST.value]).reward == env.step([api.Action.EAST.value]).reward
