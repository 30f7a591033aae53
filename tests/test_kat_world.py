"""Known-answer tests of the `World` engine, transcribed from the reference's own suites:

  [R]  src/unit_tests/test_world.rs
  [I]  tests/world_integration_tests.rs
  [P]  python/tests/test_world.py

Each test names its source.  They run against the oracle (CPU) and, with ``-m gpu``, against the
CUDA path through the N=1 ``World`` facade — same assertions, bit-exact expectations.
"""
import numpy as np
import pytest


def laser_at(world, pos):
    """[R] get_laser (test_world.rs:15-22): first listed laser at `pos` (the outermost one)."""
    for laser in world.lasers:
        if tuple(laser.pos) == tuple(pos):
            return laser
    raise AssertionError(f"No laser at {pos}")


# ----------------------------------------------------------------------------- parsing / layout
def test_tile_type(api):  # [R] test_tile_type :25
    world = api.World(
        """
    S0 . G
    L0E X @
    """
    )
    world.reset()
    assert world.start_pos[0] == (0, 0)
    assert (0, 2) in [g.pos for g in world.gems]
    source = world.source_at((1, 0))
    assert source.agent_id == 0
    assert laser_at(world, (1, 1)).agent_id == 0
    assert world.exit_pos.count((1, 1)) == 1
    assert len(world.wall_pos) == 2
    assert (1, 2) in world.wall_pos and (1, 0) in world.wall_pos


def test_duplicate_start_pos(api):  # [R] :64
    with pytest.raises(api.ParsingError, match="DuplicateStartTile"):
        api.World("S0 S0 X X")


def test_start_pos_order(api):  # [R] :78
    world = api.World("S1 S0 X X")
    assert world.start_pos == [(0, 1), (0, 0)]
    world.reset()
    assert world.agents_positions == [(0, 1), (0, 0)]


def test_start_pos_order_lvl6(api):  # [R] :89
    world = api.World.from_file("lvl6")
    world.reset()
    for agent, pos in enumerate(world.start_pos):
        assert pos == (0, agent + 4)
        assert world.agents_positions[agent] == (0, agent + 4)


def test_laser_blocked_by_wall(api):  # [R] :100
    w = api.World(
        """
        . L0S .
        .  .  .
        X  @  S0
        .  .  ."""
    )
    w.reset()
    for pos in w.laser_pos:
        assert pos != (2, 1) and pos != (3, 1)


def test_empty_world(api):  # [R] :188, [I] parse_empty_world :95
    with pytest.raises(api.ParsingError, match="EmptyWorld"):
        api.World("")


def test_parse_inconsistent_row_lengths(api):  # [R] :358
    with pytest.raises(api.ParsingError, match="InconsistentDimensions"):
        api.World(
            """X S0 .
         . ."""
        )


def test_parse_inconsistent_start_exit_tiles(api):  # [R] :381
    with pytest.raises(api.ParsingError, match="NotEnoughExitTiles"):
        api.World("S1 S0 X")


def test_parse_no_agents(api):  # [R] :395 ; [P] test_parse_wrong_worlds
    with pytest.raises(api.ParsingError, match="NoAgents"):
        api.World(". . G")
    with pytest.raises(api.ParsingError):
        api.World("X G")
    with pytest.raises(api.ParsingError):
        api.World(
            """
            @ @  @ @
            @ S0 . @
            @ .  . @
            @ @  @ @"""
        )


def test_invalid_tile(api):  # parser_v1.rs:163-169
    with pytest.raises(api.ParsingError, match="InvalidTile"):
        api.World("S0 Q X")


def test_standard_levels(api):  # [R] :439 / :449 ; [P] test_get_standard_level
    expected = {1: (1, 1), 2: (2, 1), 3: (2, 1), 4: (2, 1), 5: (4, 5), 6: (4, 4)}
    for level in range(1, 7):
        for w in (api.World.level(level), api.World.from_file(f"lvl{level}"), api.World.from_file(f"level{level}")):
            assert (w.height, w.width) == (12, 13)
            assert (w.n_agents, w.n_gems) == expected[level]


def test_laser_on_start_pos_error(api):  # [P] test_laser_on_start_pos_error ; test_parser_v1.rs:61-77
    with pytest.raises(api.ParsingError, match="AgentWithoutStart"):
        api.World(
            """
    S0  S1 X . X
    L1N .  . . .
    """
        )


def test_laser_sources_in_wall_pos(api):  # [P]
    world = api.World(
        """
        S0 . . X
       L0E . . X
        S1 . . X
       L1E . . X
"""
    )
    for source in world.laser_sources:
        assert source.pos in world.wall_pos


def test_laser_num_higher_than_n_agents(api):  # [P]
    world = api.World("S0 L1E X")
    assert world.source_at((0, 1)).agent_id == 1


def test_n_laser_colours(api):  # [P] test_n_laser_colours*
    assert api.World("S0 L0E X\nS1 L1E X").n_laser_colours == 2
    assert api.World("S0 L0E X\n . L2E X").n_laser_colours == 2
    assert api.World("S0 L0E X\n . L0E X").n_laser_colours == 1


def test_many_agents(api):  # [P]
    rows = [" .   .   . . . ."] + [f"S{k}  L{k}W  . . . X" for k in range(14)]
    world = api.World("\n".join(rows))
    assert world.n_agents == 14
    assert len(world.laser_sources) == 14


def test_source_laser_id_is_source_index(api):  # [R] check_source_laser_id_is_the_same_as_source_index :682
    w = api.World(
        """
S0  S1  S2  S3  S4  S5  S6  S7  S8  S9  S10
L0S L1S L2S L3S L4S L5S L6S L7S L8S L9S L10S
 .  .   .   .   .   .   .   .   .   .   .
 X  X   X   X   X   X   X   X   X   X   X
"""
    )
    for i, source in enumerate(w.laser_sources):
        assert source.laser_id == i


def test_beam_cells(api):  # [R] test_beam_single_source :662, test_beam_long :698, test_beam_south_direction :721
    w = api.World("L0E . L0S\nS0  X  @")
    w.reset()
    assert sorted(l.pos for l in w.lasers if l.laser_id == w.source_at((0, 0)).laser_id) == [(0, 1)]
    assert [l for l in w.lasers if l.laser_id == w.source_at((0, 2)).laser_id] == []

    w = api.World("L0E .  .  .  .  @\nS0  .  .  .  X  @")
    assert sorted(l.pos for l in w.lasers) == [(0, 1), (0, 2), (0, 3), (0, 4)]

    w = api.World("@  L0S @\n.  .   .\n.  .   .\nS0 X   @")
    assert sorted(l.pos for l in w.lasers) == [(1, 1), (2, 1), (3, 1)]


def test_beam_two_sources(api):  # [R] :757
    w = api.World(
        """
        @ L0E .  .  .
        .  .  .  .  .
        @ L1E .  @  @
        S0 .  .  X  .
        S1 .  .  X  ."""
    )
    by_id = {}
    for l in w.lasers:
        by_id.setdefault(l.laser_id, []).append(l.pos)
    assert len(by_id[w.source_at((0, 1)).laser_id]) == 3
    assert len(by_id[w.source_at((2, 1)).laser_id]) == 1


def test_laser_on_exit(api):  # [I] :482
    w = api.World(
        """
    .   L0S S1
    S0   .   .
    L1E  X   X"""
    )
    w.reset()
    lasers = w.lasers
    assert len(lasers) == 4
    for laser_id in range(2):
        assert len([l for l in lasers if l.laser_id == laser_id]) == 2


def test_laser_id(api):  # [I] test_laser_id :410, test_laser_sources_have_different_laser_ids :459
    w = api.World(
        """
        S0 .   G  X
        .  .  L0W .
        .  S1  .  X
        .  .  L0W  ."""
    )
    w.reset()
    ids = {s.pos[0]: s.laser_id for s in w.laser_sources}
    for l in w.lasers:
        assert l.laser_id == ids[l.pos[0]]
    w = api.World("L0E . L0E . X S0")
    assert len({s.laser_id for s in w.laser_sources}) == 2


# ----------------------------------------------------------------------------- stepping
def test_laser_blocked_on_reset(api):  # [R] :118
    w = api.World(
        """
        @ @ L0S @  @
        @ .  .  .  @
        @ X  S0 .  @
        @ .  .  .  @
        @ @  @  @  @"""
    )
    w.reset()
    assert all(a.is_alive for a in w.agents)
    assert laser_at(w, (1, 2)).is_on
    assert laser_at(w, (2, 2)).is_off
    assert laser_at(w, (3, 2)).is_off


FACING = """
         @ @ L0S @  @
         @ X  .  S0 @
         @ .  .  .  @
         @ X  .  S1 @
         @ @ L1N  @ @"""


def test_facing_lasers(api):  # [R] :138
    w = api.World(FACING)
    w.reset()
    w.step([api.Action.WEST, api.Action.WEST])
    assert all(a.is_alive for a in w.agents)
    assert all(l.is_off for l in w.lasers)


def test_facing_lasers_agent_dies(api):  # [R] :172
    w = api.World(FACING)
    w.reset()
    w.step([api.Action.WEST, api.Action.STAY])
    assert w.agents[0].is_dead


def test_event_exit_when_staying(api):  # [R] :158
    w = api.World("S0 X .\nS1 . X")
    w.reset()
    assert len(w.step([api.Action.EAST, api.Action.STAY])) == 1
    assert len(w.step([api.Action.STAY, api.Action.STAY])) == 0


def test_complex_laser_blocking(api):  # [R] :256
    w = api.World(
        """
    G L0E X . X
    G G . . L1W
    @ S0 . . @
    . @ . . .
    S1 G . . G"""
    )
    w.reset()
    assert laser_at(w, (0, 3)).is_on
    w.set_state(api.WorldState([(0, 2), (0, 3)], [False] * 5))
    assert laser_at(w, (0, 3)).is_off
    assert all(a.is_alive for a in w.agents)
    w.step([api.Action.STAY, api.Action.EAST])
    assert all(a.is_alive for a in w.agents)
    assert laser_at(w, (0, 3)).is_off


def test_die_in_void(api):  # [R] :324
    w = api.World("S0 V X")
    w.reset()
    w.step([api.Action.EAST])
    assert w.agents[0].is_dead


def test_num_gems_collected_and_arrived(api):  # [R] test_num_gems_collected :332, test_num_agents_arrived :345
    world = api.World("S0 G X")
    world.reset()
    assert world.gems_collected == 0
    world.step([api.Action.EAST])
    assert world.gems_collected == 1 and not world.agents[0].has_arrived
    world.step([api.Action.STAY])
    assert world.gems_collected == 1 and not world.agents[0].has_arrived
    world.step([api.Action.EAST])
    assert world.gems_collected == 1 and world.agents[0].has_arrived


def test_vertex_conflict_rust(api):  # [R] :406
    w = api.World("S0 X .\n.  . .\nS1 X .")
    w.reset()
    w.step([api.Action.SOUTH, api.Action.NORTH])
    assert w.agents_positions == [(0, 0), (2, 0)]


def test_reset_event_lists(api):  # [R] test_reset :422
    w = api.World("S0 G X")
    for _ in range(10):
        w.reset()
        assert w.agents_positions[0] == (0, 0)
        assert w.step([api.Action.EAST]) == [api.WorldEvent(api.EventType.GEM_COLLECTED, 0)]
        assert w.step([api.Action.EAST]) == [api.WorldEvent(api.EventType.AGENT_EXIT, 0)]


def test_beam_agent_on_beam_tile(api):  # [R] :738
    w = api.World(
        """
        L0E .  .  .
        .   S0 .  X
        .   .  .  . """
    )
    w.reset()
    w.step([api.Action.NORTH])
    assert w.agents_positions[0] == (0, 1)
    assert sorted(l.pos for l in w.lasers) == [(0, 1), (0, 2), (0, 3)]
    assert all(l.is_off for l in w.lasers)


def test_available_actions_rust(api):  # [I] :4
    w = api.World("S0 . G\nL0E X .")
    w.reset()
    assert sorted(w.available_actions()[0]) == sorted([api.Action.STAY, api.Action.EAST])


AVAIL_MAP = """
    .  S1 .
    .  S0 G
    L0E X X
    """


def test_available_actions_exit_trace(api):  # [I] test_available_actions_two_agents :20, test_available_actions_exit :44
    A = api.Action
    w = api.World(AVAIL_MAP)
    w.reset()

    def check(e0, e1):
        av = w.available_actions()
        assert sorted(av[0]) == sorted(e0) and sorted(av[1]) == sorted(e1)

    check([A.STAY, A.EAST, A.WEST, A.SOUTH], [A.STAY, A.EAST, A.WEST])
    w.step([A.SOUTH, A.EAST])
    check([A.STAY], [A.STAY, A.WEST, A.SOUTH])
    w.step([A.STAY, A.SOUTH])
    check([A.STAY], [A.STAY, A.WEST, A.SOUTH, A.NORTH])
    w.step([A.STAY, A.SOUTH])
    check([A.STAY], [A.STAY])


def test_available_actions_order(api):  # world.rs:349-351 : Stay first, then N, E, S, W
    A = api.Action
    w = api.World(". . .\n. S0 .\n. . X")
    w.reset()
    assert w.available_actions()[0] == [A.STAY, A.NORTH, A.EAST, A.SOUTH, A.WEST]


def test_take_action_not_available(api):  # [I] :106, :125, :144, :163
    A = api.Action
    for text, actions in [
        ("S0 X", [A.NORTH]),
        ("S0 X\nS1 X", [A.SOUTH, A.NORTH]),
        ("L0E X\nS0 .", [A.NORTH]),
        ("L0E X\nS0 .", [A.WEST]),
    ]:
        w = api.World(text)
        w.reset()
        before = w.get_state()
        with pytest.raises(api.InvalidActionError):
            w.step(actions)
        assert w.get_state() == before  # world.rs:444-453: no mutation on error


def test_wrong_number_of_actions(api):  # world.rs:436-441 -> ValueError (pyexceptions.rs)
    w = api.World("S0 X\nS1 X")
    w.reset()
    with pytest.raises(ValueError):
        w.step([api.Action.STAY])


def test_dead_agent_does_not_block_the_laser(api):  # [I] :279 — two-pass death cascade
    A, E = api.Action, api.EventType
    w = api.World(
        """
        S0 .   G  X
        .  .  L2W X
        .  S1  .  X
        . L1N  .  S2"""
    )
    w.reset()
    events = w.step([A.EAST, A.NORTH, A.STAY])
    # pass 1: agent 1 dies in L2W; pass 2: the released L1N beam kills agent 0 (world.rs:468-472)
    assert events == [api.WorldEvent(E.AGENT_DIED, 1), api.WorldEvent(E.AGENT_DIED, 0)]
    for l in w.lasers:
        if l.pos == (0, 1):
            assert l.is_on
    # the cell between the dead blocker and the next leaver keeps a stale off bit
    l1n = {l.pos: l.is_on for l in w.lasers if l.laser_id == 1}
    assert l1n == {(2, 1): True, (1, 1): False, (0, 1): True}


def test_world_state_equal(api):  # [I] :312
    w = api.World("S0 . G\n.  . X")
    w.reset()
    s1 = w.get_state()
    assert s1 == w.get_state()
    w.step([api.Action.STAY])
    assert s1 == w.get_state()
    w.step([api.Action.EAST])
    assert s1 != w.get_state()


def test_blocked_laser_on_spawn(api):  # [I] :559
    w = api.World(
        """
    . L1S .  X .  .
    . S1  .  . .  .
    . S0  .  @ . L0W
    .  .  .  . .  .
    @  . L2N . .  X
    .  .  .  @ .  . """
    )
    actions = [api.Action.EAST, api.Action.EAST]
    w.reset()
    w.step(actions)
    w.reset()
    w.step(actions)


def test_reset_in_blocked_laser(api):  # [P]
    A = api.Action
    w = api.World(
        """
    . .  .  @ .  .  .  . .  .  . . .
    . .  .  . @  X  .  . .  @  . . .
    . .  .  . . L1S .  X .  .  @ . @
    . .  .  . .  .  .  . .  .  . . .
    . S3  . . . S1  .  . .  .  . @ .
    . .  .  @ . S0  .  @ . L0W . . .
    @ .  .  @ .  .  .  . .  .  . . .
    . .  .  . .  .  .  . .  .  . X .
    . .  .  . @  . L2N . .  .  . . .
    . .  S2 . .  .  .  . .  .  . . X
    . .  .  @ .  .  @  . .  @  . . .
    . .  .  . .  .  .  @ .  .  . . ."""
    )
    w.reset()
    actions = [A.EAST, A.EAST, A.NORTH, A.SOUTH]
    ev1 = w.step(actions)
    w.reset()
    assert w.step(actions) == ev1


def test_world_step_one_action(api):  # [P]
    world = api.World("S0 X . .\n.  . . .\n.  . . .")
    world.reset()
    assert world.step(api.Action.SOUTH) == []
    assert world.agents_positions == [(1, 0)]


def test_world_step_type_errors(api):  # [P] test_world_step_something_else_than_action, ..._invalid_sequence_action
    world = api.World("S0 X . .\n.  . . .\n.  . . .")
    world.reset()
    with pytest.raises(TypeError):
        world.step(23)
    world.step((api.Action.SOUTH,))
    assert world.agents_positions == [(1, 0)]
    with pytest.raises(TypeError, match="Action must be of type Action or list\\[Action\\]"):
        world.step((23,))


def test_walk_into_wall(api):  # [P]
    world = api.World(
        """@ @ @  @ @ @
@ . .  . . @
@ . S0 . . @
@ . .  X . @
@ @ @  @ @ @"""
    )
    world.reset()
    world.step([api.Action.SOUTH])
    with pytest.raises(api.InvalidActionError):
        world.step([api.Action.SOUTH])


def test_gem_collected_and_agent_died(api):  # [P] — quirk 5: a gem under a lethal beam is not collected
    world = api.World("S0  G  X\nS1 L1N X")
    world.reset()
    events = world.step([api.Action.EAST, api.Action.STAY])
    assert len(events) == 1
    assert events[0].event_type == api.EventType.AGENT_DIED
    assert world.gems_collected == 0


def test_world_gem_collected_and_agent_has_arrived(api):  # [P]
    A = api.Action
    world = api.World("S0 X . .\n.  . . .\nG  . . .")
    world.reset()
    world.reset()
    world.step([A.SOUTH])
    world.step([A.SOUTH])
    assert world.gems_collected == 1
    world.step([A.NORTH])
    world.step([A.NORTH])
    world.step([A.EAST])
    assert world.agents[0].has_arrived


def test_vertex_conflict_python(api):  # [P]
    world = api.World(
        """
        .  X  .  .
        S0 .  S1  .
        .  X  .  ."""
    )
    world.reset()
    state = world.get_state()
    world.step([api.Action.EAST, api.Action.WEST])
    assert state == world.get_state()


def test_swapping_conflict(api):  # [P]
    A = api.Action
    world = api.World("S0 X  .  .\n.  .  S1  .\n.  X  .  .")
    world.reset()
    world.step([A.SOUTH, A.WEST])
    with pytest.raises(api.InvalidActionError):
        world.step([A.EAST, A.WEST])


def test_walk_into_laser_source(api):  # [P]
    A = api.Action
    world = api.World(
        """
        @ L0S @
        .  .  .
        X  .  S0
        .  .  ."""
    )
    world.reset()
    world.step([A.WEST])
    world.step([A.NORTH])
    with pytest.raises(ValueError):
        world.step([A.NORTH])


def test_walk_outside_map(api):  # [P]
    A = api.Action
    world = api.World(
        """@ @ L0S @  @
@ .  .  .  @
@ X  .  S0 @
@ .  .  .  @
@ @  .  @  @
"""
    )
    world.reset()
    world.step([A.SOUTH])
    world.step([A.WEST])
    world.step([A.SOUTH])
    with pytest.raises(ValueError):
        world.step([A.SOUTH])


def test_world_done(api):  # [P] — a dead agent may only STAY
    A = api.Action
    world = api.World(
        """
G  G  . .  S1
X  .  . @  .
@  .  G .  .
G  .  . G  X
@ L0N . S0 ."""
    )
    world.reset()
    for _ in range(3):
        world.step([A.STAY, A.WEST])
    assert world.agents[1].is_dead
    with pytest.raises(ValueError):
        world.step([A.STAY, A.WEST])


def test_laser_tile_state(api):  # [P]
    world = api.World("L0E S0 . X")
    world.reset()
    assert len(world.lasers) == 3
    assert all(l.is_off for l in world.lasers)
    world.step([api.Action.EAST])
    for laser in world.lasers:
        assert laser.is_on == (laser.pos == (0, 1))


def test_no_reset(api):  # [P] World::new resets itself (world.rs:82)
    w = api.World("S0 . X")
    w.step(api.Action.EAST)
    assert w.agents_positions == [(0, 1)]


def test_available_actions_python(api):  # [P] test_available_actions
    A = api.Action
    world = api.World(
        """
@ @ L0S @  @
@ .  .  .  @
@ X  .  S0 @
@ X  .  S1 @
@ @  @  @  @
"""
    )
    world.reset()
    av = world.available_actions()
    assert sorted(av[0]) == sorted([A.NORTH, A.WEST, A.STAY])
    assert sorted(av[1]) == sorted([A.WEST, A.STAY])


# ----------------------------------------------------------------------------- get_state / set_state
SMALL = """
        S0 . G
        X  . .
    """


def test_force_state_invalid_sizes(api):  # [R] :199, :228
    w = api.World(SMALL)
    w.reset()
    with pytest.raises(api.InvalidWorldStateError, match="InvalidNumberOfAgents"):
        w.set_state(api.WorldState([(1, 2), (0, 0)], [True]))
    with pytest.raises(api.InvalidWorldStateError, match="InvalidNumberOfGems"):
        w.set_state(api.WorldState([(1, 2)], [True, False]))


def test_set_state_available_actions(api):  # [R] :303
    A = api.Action
    w = api.World(
        """
        .  . . @ . . . @ . X
        .  @ . @ . @ . @ . @
        S0 @ . . . @ . . . @
    """
    )
    w.reset()
    w.set_state(api.WorldState([(0, 0)], []))
    assert sorted(w.available_actions()[0]) == sorted([A.SOUTH, A.STAY, A.EAST])


def test_force_state(api):  # [R] test_force_state :456, test_force_end_state :473
    w = api.World(SMALL)
    w.reset()
    w.set_state(api.WorldState([(1, 2)], [True]))
    assert w.agents_positions[0] == (1, 2)
    assert w.gems[0].is_collected
    w.set_state(api.WorldState([(1, 0)], [True]))
    assert w.agents_positions[0] == (1, 0)
    assert w.gems[0].is_collected


def test_force_state_agent_dies(api):  # [R] :490 ; [I] :230
    w = api.World(
        """
        S0 S1 G
        X  X L0W
    """
    )
    w.reset()
    w.set_state(api.WorldState([(1, 0), (1, 1)], [False], [True, False]))
    assert w.agents[0].has_arrived
    assert not w.agents[1].has_arrived
    assert w.agents[1].is_dead


def test_wrong_world_state(api):  # [R] :527 ; [P] test_set_invalid_state_dead
    w = api.World(
        """
        S0 L0S X
        S1  .  X
    """
    )
    w.reset()
    with pytest.raises(api.InvalidWorldStateError, match="InvalidWorldState"):
        w.set_state(api.WorldState([(0, 0), (1, 1)], []))
    w = api.World("S0 L0S X\nS1  .  X")
    with pytest.raises(api.InvalidWorldStateError):
        w.set_state(api.WorldState([(0, 0), (0, 1)], [], [True, True]))


def test_force_state_agents_have_exited(api):  # [I] :182
    w = api.World(SMALL)
    w.reset()
    events = w.set_state(api.WorldState([(1, 0)], [True]))
    assert all(a.has_arrived for a in w.agents)
    assert events == [api.WorldEvent(api.EventType.AGENT_EXIT, 0)]


def test_force_wrong_state_check_laser_not_blocked(api):  # [I] :209
    w = api.World(
        """
        S1  S0 X
        L0E  G  X
    """
    )
    w.reset()
    with pytest.raises(api.InvalidWorldStateError, match="InvalidAgentPosition"):
        w.set_state(api.WorldState([(1, 1), (1, 0)], [True]))
    assert all(l.is_on for l in w.lasers)
    assert all(not g.is_collected for g in w.gems)


def test_set_invalid_state_restores_positions(api):  # [I] test_set_invalid_state :254
    w = api.World(
        """
        S0 S1 X
        @  @  X
    """
    )
    w.reset()
    with pytest.raises(api.InvalidWorldStateError, match="InvalidAgentPosition"):
        w.set_state(api.WorldState([(1, 0), (1, 1)], []))
    assert w.agents_positions == [(0, 0), (0, 1)]


def test_world_state_dead_agents(api):  # [I] :538
    w = api.World(
        """
    S0 . G
    V  . X
    """
    )
    w.reset()
    assert w.get_state().agents_alive[0]
    w.step([api.Action.SOUTH])
    state = w.get_state()
    assert not state.agents_alive[0]
    w.reset()
    w.set_state(state)
    assert not w.agents[0].is_alive


def test_get_state(api):  # [P]
    world = api.World("S0 G X")
    world.reset()
    state = world.get_state()
    assert state.agents_positions == [(0, 0)] and state.gems_collected == [False]
    world.step([api.Action.EAST])
    state = world.get_state()
    assert state.agents_positions == [(0, 1)] and state.gems_collected == [True]


def test_set_state(api):  # [P]
    world = api.World("S0 G X")
    world.reset()
    world.step([api.Action.EAST])
    events = world.set_state(api.WorldState([(0, 0)], [False]))
    assert world.agents_positions == [(0, 0)]
    assert world.gems_collected == 0
    assert len(events) == 0
    events = world.set_state(api.WorldState([(0, 2)], [True]))
    assert world.agents_positions == [(0, 2)]
    assert world.gems_collected == 1
    assert len(events) == 1
    assert events[0].agent_id == 0
    assert events[0].event_type == api.EventType.AGENT_EXIT


def test_set_invalid_state(api):  # [P]
    world = api.World(
        """
        S1  S0 X
        L0E  G  X"""
    )
    world.reset()
    with pytest.raises(api.InvalidWorldStateError):
        world.set_state(api.WorldState([(0, 0), (0, 1)], [True, True]))
    with pytest.raises(api.InvalidWorldStateError):
        world.set_state(api.WorldState([(0, 0)], [True]))
    with pytest.raises(IndexError):
        world.set_state(api.WorldState([(10, 1), (1, 0)], [True]))
    with pytest.raises(api.InvalidWorldStateError):
        world.set_state(api.WorldState([(1, 1), (1, 0)], [True]))
    with pytest.raises(api.InvalidWorldStateError):
        world.set_state(api.WorldState([(0, 0), (0, 0)], [True]))


def test_set_state_agent_dead(api):  # [P]
    world = api.World("S0 G X")
    world.reset()
    world.set_state(api.WorldState([(0, 0)], [False], [False]))
    assert not world.agents[0].is_alive


def test_world_state_hash_eq(api):  # [P] test_world_state_hash_eq, _dead, _neq
    world = api.World("S0 G X")
    world.reset()
    s1, s2 = world.get_state(), world.get_state()
    assert hash(s1) == hash(s2) and s1 == s2
    a = api.WorldState([(0, 0)], [False], [True])
    b = api.WorldState([(0, 0)], [False], [False])
    assert a != b and hash(a) != hash(b)
    assert api.WorldState([(0, 0)], [False]) != api.WorldState([(0, 1)], [False])


def test_world_state_constructor(api):  # [P]
    assert all(api.WorldState([(0, 0)], [False]).agents_alive)
    assert api.WorldState([(0, 0), (1, 1)], [True], [False, True]).agents_alive == [False, True]


def test_state_from_to_array(api):  # [P] — exact vectors (pyworld_state.rs:79-132)
    s = api.WorldState([(0, 0)], [False])
    assert list(s.as_array()) == [0.0, 0.0, 0.0, 1.0]
    assert api.WorldState.from_array([0.0, 0.0, 0.0, 1.0], 1, 1) == s
    s = api.WorldState([(25, 17), (10, 30)], [True, False], agents_alive=[True, False])
    expected = [25.0, 17.0, 10.0, 30.0, 1.0, 0.0, 1.0, 0.0]
    assert list(s.as_array()) == expected
    assert s.as_array().dtype.name == "float32"
    assert api.WorldState.from_array(expected, 2, 2) == s
    with pytest.raises(ValueError):
        api.WorldState.from_array(expected, 2, 1)


def test_world_n_agents(api):  # [P]
    assert api.World("S0 S1 X X").n_agents == 2
    assert api.World.level(6).n_agents == 4


def test_world_tiles(api):  # [P]
    w = api.World("S0 . X")
    assert w.start_pos == [(0, 0)]
    assert w.random_start_pos == [[(0, 0)]]
    assert w.exit_pos == [(0, 2)]


# ----------------------------------------------------------------------------- laser-source mutators
# LaserBeam::{set_agent_id, enable, disable} (src/core/tiles/laser.rs:69-84) through PyLaserSource
# (src/bindings/tiles/pylaser_source.rs:55-142).  [I] tests/world_integration_tests.rs, [P] python/tests/test_world.py
SRC_MAP = "S0 .   G  X\n.  .  L0W .\n.  S1  .  X\n.  .   .  ."


def test_change_laser_id(api):  # [I] change_laser_id :333
    w = api.World(SRC_MAP)
    w.reset()
    assert all(l.agent_id == 0 for l in w.lasers)
    source = w.laser_sources[0]
    source.agent_id = 1
    assert source.agent_id == 1 and w.laser_sources[0].agent_id == 1
    assert all(l.agent_id == 1 for l in w.lasers)
    events = w.step([api.Action.SOUTH, api.Action.STAY])  # agent 0 dies in what is now agent 1's laser
    assert events == [api.WorldEvent(api.EventType.AGENT_DIED, 0)]


def test_disable_laser_source(api):  # [I] disable_laser_source :363
    w = api.World(SRC_MAP)
    w.reset()
    assert all(l.is_on for l in w.lasers)
    source = w.laser_sources[0]
    source.disable()
    assert all(l.is_off for l in w.lasers) and not any(l.is_enabled for l in w.lasers)
    assert w.laser_sources[0].is_disabled
    source.enable()
    assert all(l.is_on for l in w.lasers)


def test_disable_laser_source_and_block_with_agent(api):  # [I] :384
    w = api.World("L0E . S0 X")
    w.reset()
    assert laser_at(w, (0, 1)).is_on
    w.laser_sources[0].disable()
    assert laser_at(w, (0, 1)).is_off
    w.step([api.Action.WEST])
    assert laser_at(w, (0, 2)).is_off
    w.step([api.Action.EAST])
    assert laser_at(w, (0, 1)).is_off


def test_disable_laser_then_reset_does_not_turn_on(api):  # [I] :447
    w = api.World("L0E . S0 X")
    w.reset()
    w.laser_sources[0].disable()
    w.reset()
    laser = laser_at(w, (0, 1))
    assert not laser.is_enabled and laser.is_off


def test_enable_turns_the_whole_beam_on(api):  # laser.rs:69-72: enable() is turn_on(0), whoever stands in the beam
    w = api.World("L0E . S0 X")
    w.reset()
    w.step([api.Action.WEST])  # the owner blocks its beam at offset 0
    assert laser_at(w, (0, 1)).is_off and laser_at(w, (0, 2)).is_off
    source = w.laser_sources[0]
    source.is_enabled = False
    source.is_disabled = False  # assignment forms of disable() / enable() (pylaser_source.rs:84-92)
    assert laser_at(w, (0, 1)).is_on and laser_at(w, (0, 2)).is_on
    w.step([api.Action.STAY])  # leave + pre_enter on the same cell cut it again (laser.rs:173-202)
    assert laser_at(w, (0, 1)).is_off


def test_disable_deadly_laser_source_and_walk_into_it(api):  # [P] :470
    world = api.World("L0S . L0W X\nS0 S1  .  X")
    world.reset()
    world.source_at((0, 2)).disable()
    events = world.step([api.Action.STAY, api.Action.NORTH])
    assert len(events) == 0
    assert all(a.is_alive for a in world.agents)


def test_change_laser_colour(api):  # [P] :485
    world = api.World("L1E . S1 S0 X\nL0E .  .  . X")
    world.reset()
    assert len(world.lasers) == 8
    for laser in world.lasers:
        assert laser.agent_id == (1 if laser.pos[0] == 0 else 0)
    bot_source = world.source_at((1, 0))
    bot_source.set_colour(1)
    world.reset()
    for laser in world.lasers:
        if laser.pos[0] == 1:
            assert laser.agent_id == 1
    events = world.step([api.Action.SOUTH, api.Action.SOUTH])
    assert len(events) == 0
    assert all(a.is_alive for a in world.agents)


def test_change_laser_colour_errors(api):  # [P] :517-575
    world = api.World("L0E S0 . X")
    world.reset()
    source = world.source_at((0, 0))
    with pytest.raises(OverflowError):
        source.set_colour(-1)
    for bad in (2, 1):  # there is only one agent
        with pytest.raises(ValueError):
            source.set_colour(bad)
        with pytest.raises(ValueError):
            source.agent_id = bad
    world = api.World("L0E X X S0 S1")  # agent 1 would be killed on reset
    world.reset()
    with pytest.raises(ValueError):
        world.source_at((0, 0)).agent_id = 1


def test_laser_colour_change_remains_after_reset(api):  # [P] :527
    world = api.World("L0E X X @ S0 S1")
    world.reset()
    world.source_at((0, 0)).agent_id = 1
    world.reset()
    assert world.source_at((0, 0)).agent_id == 1


def test_change_laser_colour_back(api):  # [P] :578
    world = api.World("L1E . S1 S0 X\nL0E .  .  . X")
    world.reset()
    bot_source = world.source_at((1, 0))
    bot_source.set_colour(1)
    world.reset()
    assert world.source_at((1, 0)).agent_id == 1
    assert all(l.agent_id == 1 for l in world.lasers)
    bot_source.set_colour(0)
    world.reset()
    assert world.source_at((1, 0)).agent_id == 0
    for laser in world.lasers:
        assert laser.agent_id == (1 if laser.pos[0] == 0 else 0)
    assert world.n_laser_colours == 2


# ----------------------------------------------------------------------------- World.exit_pos setter
def test_set_exit_positions(api):  # [P] test_set_exit_positions :764 ; world.rs:195-234
    A, E = api.Action, api.EventType
    world = api.World("S0 . X")
    world.reset()
    assert world.exit_pos[0] == (0, 2)
    world.exit_pos = [(0, 1)]
    world.reset()
    assert world.exit_pos[0] == (0, 1)
    events = world.step([A.EAST])
    assert events[0].event_type == E.AGENT_EXIT
    world.exit_pos = [(0, 2)]
    world.reset()
    assert world.exit_pos[0] == (0, 2)
    assert len(world.step([A.EAST])) == 0
    events = world.step([A.EAST])
    assert events[0].event_type == E.AGENT_EXIT


def test_observe_layered_change_exits(api):  # [O] python/tests/test_observations.py:76
    world = api.World("S0 X . .")
    assert world.exit_pos[0] == (0, 1) and len(world.exit_pos) == 1
    world.exit_pos = [(0, 2), (0, 3)]
    world.reset()
    obs = world.observe_layered()
    EXIT = 2 * 1 + 3
    assert np.all(obs[:, EXIT, 0, 2] == 1) and np.all(obs[:, EXIT, 0, 3] == 1)
    assert obs[:, EXIT].sum() == 2


def test_set_exit_positions_errors(api):  # world.rs:196-201 ; the reference panics on non-floor tiles (:216-229)
    world = api.World("S0 . X\nS1 @ X")
    with pytest.raises(api.ParsingError, match="NotEnoughExitTiles"):
        world.exit_pos = [(0, 1)]
    with pytest.raises(Exception):  # the reference panics half-way (and poisons the world's mutex); here the call is refused
        world.exit_pos = [(0, 1), (1, 1)]  # a wall


def test_exit_under_a_laser_moves(api):  # world.rs:208-213 / :222-227: the tile under the (single) laser is replaced
    world = api.World("L0E . X\nS0  . .\nS1  . X")
    world.reset()
    world.exit_pos = [(0, 1), (2, 2)]
    world.reset()
    assert laser_at(world, (0, 1)).is_on and laser_at(world, (0, 2)).is_on
    world.step([api.Action.EAST, api.Action.STAY])
    events = world.step([api.Action.NORTH, api.Action.STAY])  # agent 0 enters its own beam on the new exit
    assert events == [api.WorldEvent(api.EventType.AGENT_EXIT, 0)]
    assert laser_at(world, (0, 2)).is_off


# ---- copies, pickling, joint actions (the rest of the PyWorld surface)
def test_deepcopy(api):  # [P] python/tests/test_world.py:300
    from copy import deepcopy

    world = api.World("S0 . X")
    world2 = deepcopy(world)
    assert world.agents_positions == world2.agents_positions
    assert world.agents_positions is not world2.agents_positions
    assert world.width == world2.width


def test_deepcopy_not_initial_state(api):  # [P] python/tests/test_world.py:308
    from copy import deepcopy

    world = api.World("S0 . X")
    world.reset()
    world.step([api.Action.EAST])
    world2 = deepcopy(world)
    assert world.agents_positions == world2.agents_positions == [(0, 1)]
    assert world.width == world2.width
    world2.step([api.Action.EAST])  # world.rs:645-652: the clone is a world of its own
    assert world.agents_positions == [(0, 1)] and world2.agents_positions == [(0, 2)]
    assert world.get_state() != world2.get_state()


def test_pickle_world_state(api):  # [P] python/tests/test_serialization.py:8
    import pickle
    import random

    rng = random.Random(0)
    for _ in range(50):
        s = api.WorldState(gems_collected=[rng.choice([True, False]) for _ in range(rng.randint(0, 10))],
                           agents_positions=[(rng.randint(0, 50), rng.randint(0, 90)) for _ in range(rng.randint(0, 10))])
        assert pickle.loads(pickle.dumps(s)) == s


def test_pickle_world(api):  # [P] python/tests/test_serialization.py:19 (5 steps per level instead of 20)
    import pickle
    import random

    rng = random.Random(1)
    for lvl in range(1, 7):
        world = api.World.level(lvl)
        world.reset()
        for _ in range(5):
            world.step([rng.choice(a) for a in world.available_actions()])
            other = pickle.loads(pickle.dumps(world))
            assert (other.n_agents, other.n_gems, other.height, other.width) == (world.n_agents, world.n_gems, world.height, world.width)
            assert other.exit_pos == world.exit_pos and other.start_pos == world.start_pos
            assert other.wall_pos == world.wall_pos and other.void_pos == world.void_pos
            assert world.get_state() == other.get_state()
            assert world.available_actions() == other.available_actions()


def test_pickled_world_keeps_same_laser_ids(api):  # [P] python/tests/test_serialization.py:41
    import pickle

    world = api.World("L0E L1S S0 S1 X X")
    other = pickle.loads(pickle.dumps(world))
    for source in world.laser_sources:
        assert source in other.laser_sources
        mine, theirs = world.source_at(source.pos), other.source_at(source.pos)
        assert (mine.laser_id, mine.agent_id, mine.direction) == (theirs.laser_id, theirs.agent_id, theirs.direction)


def test_copy_keeps_mutated_sources(api):  # world.rs:98-110: get_config() reads the sources as they are now
    from copy import deepcopy

    world = api.World("L0E . .\nS0 . X\nS1 . X")
    world.reset()
    world.source_at((0, 0)).agent_id = 1
    world.source_at((0, 0)).disable()
    other = deepcopy(world)
    src = other.source_at((0, 0))
    assert src.agent_id == 1 and not src.is_enabled


def test_available_joint_actions(api):  # [D] src/bindings/world/pyworld.rs:483-485 (doc example) ; world.rs:257-263
    world = api.World(". .  .  . .\n. S0 . S1 .\n. X  .  X .\n")
    world.reset()
    joint = world.available_joint_actions()
    assert len(joint) == len(api.Action.variants()) ** 2
    assert joint[0] == [a[0] for a in world.available_actions()] and all(len(j) == 2 for j in joint)
    walled = api.World("@ @ @\n@ S0 X")
    walled.reset()
    assert walled.available_joint_actions() == [[api.Action.STAY], [api.Action.EAST]]


def test_action_from_delta_pickle_deepcopy(api):  # [P] python/tests/test_actions.py:40-90
    import copy
    import pickle

    A = api.Action
    for a in A.variants():
        assert copy.deepcopy(a) == a and pickle.loads(pickle.dumps(a)) == a
    assert A.from_delta(0, 0) == A.STAY and A.from_delta(0, -1) == A.NORTH and A.from_delta(0, 1) == A.SOUTH
    assert A.from_delta(1, 0) == A.EAST and A.from_delta(-1, 0) == A.WEST
    with pytest.raises(ValueError):
        A.from_delta(1, 1)


def test_set_agent_position_doc_example(api):  # [D] src/bindings/world/pyworld.rs:274-281
    world = api.World("S0 . . X")
    world.reset()
    assert world.set_agent_position(0, (0, 2)) == []
    events = world.step([api.Action.EAST])
    assert events[0].event_type == api.EventType.AGENT_EXIT
    with pytest.raises(ValueError):
        world.set_agent_position(1, (0, 0))
    with pytest.raises(IndexError):
        world.set_agent_position(0, (0, 9))


def test_set_agents_positions(api):  # pyworld.rs:252-263: the current state with new positions, through set_state
    world = api.World("S0 G . X\nS1 . . X")
    world.reset()
    events = world.set_agents_positions([(0, 2), (1, 3)])
    assert events == [api.WorldEvent(api.EventType.AGENT_EXIT, 1)]
    assert world.agents_positions == [(0, 2), (1, 3)] and world.get_state().gems_collected == [False]
    # onto the gem: the state handed to set_state still says "not collected", so its final comparison fails (world.rs:588-594)
    with pytest.raises(api.InvalidWorldStateError):
        world.set_agents_positions([(0, 1), (1, 3)])
    with pytest.raises(api.InvalidWorldStateError):
        world.set_agents_positions([(0, 0)])
    with pytest.raises(api.InvalidWorldStateError):
        world.set_agents_positions([(0, 2), (0, 2)])


def test_gem_at_doc_example(api):  # [D] src/bindings/world/pyworld.rs:306-314
    world = api.World("S0 G X")
    world.reset()
    assert not world.gem_at((0, 1)).is_collected
    world.step([api.Action.EAST])
    assert world.gem_at((0, 1)).is_collected
    with pytest.raises(ValueError):
        world.gem_at((0, 0))
    with pytest.raises(IndexError):
        world.gem_at((3, 0))
    wrapped = api.World("L0E G X\nS0 . .")  # the gem sits under a laser tile: Tile::Laser, not Tile::Gem (pyworld.rs:321-326)
    wrapped.reset()
    with pytest.raises(ValueError):
        wrapped.gem_at((0, 1))


def test_save_writes_the_world_string(api, tmp_path):  # pyworld.rs:183-189
    world = api.World("S0 G X")
    path = tmp_path / "map.txt"
    world.save(str(path))
    again = api.World(path.read_text())
    assert (again.width, again.height, again.n_gems) == (3, 1, 1)
    with pytest.raises(ValueError):
        world.save(str(tmp_path / "missing_dir" / "map.txt"))


# ---- small value types (python/tests/test_direction.py, test_actions.py, test_death_strategy.py, test_other.py)
def test_direction(api):  # [P] python/tests/test_direction.py:4-36
    D = api.Direction
    dirs = [D.NORTH, D.SOUTH, D.EAST, D.WEST]
    for d in dirs:
        assert d == d and all(d != d2 for d2 in dirs if d is not d2)
    assert D("N") == D.NORTH and D("W") == D.WEST
    with pytest.raises(ValueError):
        D("z")
    assert (D.NORTH.delta, D.SOUTH.delta, D.EAST.delta, D.WEST.delta) == ((-1, 0), (1, 0), (0, 1), (0, -1))
    assert D.NORTH.opposite() == D.SOUTH and D.SOUTH.opposite() == D.NORTH
    assert D.EAST.opposite() == D.WEST and D.WEST.opposite() == D.EAST


def test_action_value_type(api):  # [P] python/tests/test_actions.py:4-38
    A = api.Action
    assert A.NORTH == A.NORTH and A(0) == A(0)
    values = [A.NORTH, A.SOUTH, A.EAST, A.WEST, A.WEST]
    assert [values.count(a) for a in (A.NORTH, A.SOUTH, A.EAST, A.WEST)] == [1, 1, 1, 2]
    assert {a.name for a in A.variants()} == {"NORTH", "SOUTH", "EAST", "WEST", "STAY"}
    assert len({hash(a) for a in A.variants()}) == 5 and len(set(A.variants())) == 5


def test_end_strategy(api):  # [P] python/tests/test_death_strategy.py:4-18
    env = api.LLE("S0  G  X\nS1 L1N X")
    env.reset()
    assert env.step([api.Action.EAST.value, api.Action.STAY.value]).done


# ---- tile handles: PyGem.collect / .agent, PyLaser.agent (src/bindings/tiles/pygem.rs:51-76, pylaser.rs:61-81)
def test_gem_collect_and_tile_agents(api):
    world = api.World("S0 G . X\nS1 . . X")
    world.reset()
    gem = world.gems[0]
    assert not gem.is_collected and gem.agent is None
    gem.collect()                                     # no event, no step: the flag alone changes (gem.rs:17-19)
    assert gem.is_collected and world.gems[0].is_collected and world.gems_collected == 1
    assert world.get_state().gems_collected == [True]
    assert world.step([api.Action.EAST, api.Action.STAY]) == []   # entering a collected gem collects nothing
    assert world.gems[0].agent == 0
    world.reset()
    assert not world.gems[0].is_collected


def test_laser_tile_agent_and_wrapped_gem(api):
    world = api.World("L1E G . X\nS0 S1 . X")   # the gem sits under agent 1's beam: a Tile::Laser wraps it
    world.reset()
    lasers = {l.pos: l for l in world.lasers}
    assert lasers[(0, 1)].agent is None and not lasers[(0, 1)].is_disabled
    with pytest.raises(ValueError, match="is not a gem"):
        world.gems[0].collect()                       # pygem.rs:54-62
    world.step([api.Action.STAY, api.Action.NORTH])   # agent 1 walks into its own beam, onto the gem
    lasers = {l.pos: l for l in world.lasers}
    assert lasers[(0, 1)].agent == 1 and lasers[(0, 2)].agent is None and lasers[(0, 2)].is_off
    assert world.gems[0].is_collected and world.gems[0].agent is None   # PyGem.agent answers for top-level gems only


# ---- remaining cases of python/tests/test_world.py
def test_world_move_and_tuple_actions(api):  # [P] test_world_move, test_world_step_tuple_and_invalid_sequence_action
    world = api.World("S0 X . .\n.  . . .\n.  . . .")
    world.reset()
    world.step([api.Action.SOUTH])
    world.step([api.Action.EAST])
    world.step([api.Action.NORTH])
    assert world.agents_positions == [(0, 1)]
    world.reset()
    world.step((api.Action.SOUTH,))
    assert world.agents_positions == [(1, 0)]
    with pytest.raises(TypeError, match="Action must be of type Action or list\\[Action\\]"):
        world.step((23,))


def test_world_agents_alive(api):  # [P] test_world_agents
    world = api.World("S0 S1 S2\nX  X  X")
    world.reset()
    assert not any(a.is_dead for a in world.agents) and all(a.is_alive for a in world.agents)


def test_world_state_hash(api):  # [P] test_world_state_hash_eq_dead, test_world_state_hash_neq
    WS = api.WorldState
    s1, s2 = WS([(0, 0)], [False], [True]), WS([(0, 0)], [False], [False])
    assert hash(s1) != hash(s2) and s1 != s2
    s1, s2 = WS([(0, 0)], [False]), WS([(0, 1)], [False])
    assert hash(s1) != hash(s2) and s1 != s2


def test_set_wrong_agent_position(api):  # [P]
    world = api.World("S0 . . X")
    with pytest.raises(ValueError):
        world.set_agent_position(25, (0, 0))
    with pytest.raises(IndexError):
        world.set_agent_position(0, (0, 25))


def test_set_agents_positions_two_agents(api):  # [P]
    world = api.World("S0 . . X\nS1 . . X")
    for j in range(world.width):
        world.set_agents_positions([(0, j), (1, j)])
        assert world.agents_positions == [(0, j), (1, j)]
        world.set_agents_positions([(1, j), (0, j)])
        assert world.agents_positions == [(1, j), (0, j)]


def test_set_conflicting_agents_positions(api):  # [P]
    world = api.World("S0 . . X\nS1 . . X")
    with pytest.raises(api.InvalidWorldStateError):
        world.set_agents_positions([(0, 0), (0, 0)])


def test_change_laser_colour_errors(api):  # [P] test_change_laser_colour_to_negative_colour / _to_invalid_colour / _kills_agent_on_start
    world = api.World("L0E S0 . X")
    world.reset()
    source = world.source_at((0, 0))
    with pytest.raises(OverflowError):
        source.set_colour(-1)
    for colour in (2, 1):
        with pytest.raises(ValueError):
            source.set_colour(colour)
        with pytest.raises(ValueError):
            source.agent_id = colour
    world = api.World("L0E X X S0 S1")
    world.reset()
    with pytest.raises(ValueError):
        world.source_at((0, 0)).agent_id = 1  # agent 0 would be killed on reset


def test_laser_on_start_pos_removed(api):  # [P]
    world = api.World('world_string = """\n .  S1 X . X\nL1N .  . . ."""\n\n[[agents]]\nstart_positions = [{ i = 0, j = 0 }, { i = 1, j = 1 }]\n')
    assert world.random_start_pos[0] == [(1, 1)]  # (0, 0) would kill agent 0 in agent 1's beam on start


def test_n_laser_colours(api):  # [P] test_n_laser_colours_1agent, test_n_laser_colours_same_colours
    assert api.World("S0 L0E X\n . L2E X").n_laser_colours == 2
    assert api.World("S0 L0E X\n . L0E X").n_laser_colours == 1
