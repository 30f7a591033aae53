"""TOML v2 maps (src/core/parsing/toml/*.rs): known-answer tests transcribed from src/unit_tests/test_toml_config.rs [T] and
python/tests/test_world.py [P], run against the oracle (tomllib + restated serde models) and against the product's own
TOML reader and map compiler (CPU, no kernel), plus the start sampler on the device (-m gpu).

The start sampler's random stream is PARITY UNPINNED against the reference (rand::StdRng, no lockfile; the reference only
tests determinism per seed): the device library defines its own Philox contract (include/lle_b200.h, lle_vec_reset), which
the oracle restates.  Everything else on this page is pinned by the reference's tests."""
import numpy as np
import pytest

from oracle import lle_oracle as lo

TEST_OK = '''
width = 10
height = 5
exits = [{ j_min = 9 }]
gems = [{ i = 0, j = 2 }]
world_string = """
X . . . S1 . . . . .
. . . . .  . . . . .
. . . . .  . . . . .
. . . . .  . . . . .
. . . . .  . . . . .
"""

[[agents]]
start_positions = [{ i_min = 0, i_max = 0 }]

[[agents]]
# Deduced from the string map that agent 1 has a start position at (0, 5).

[[agents]]
start_positions = [{ i = 0, j = 5 }, { i = 3, j = 5 }]

[[agents]]
start_positions = [
    { i = 4, j = 9 },
    { i_min = 1, i_max = 3, j_min = 0, j_max = 3 },
    { j_min = 4 },
]
'''
GLOBAL_STARTS = "width = 10\nheight = 10\nn_agents = 5\nstarts = [{ row = 0 }]\nexits = [{ col = 4 }]\n"
ANYWHERE = "width = 10\nheight = 10\nexits = [{ i = 0, j = 9 }]\n\n[[agents]]\n# Start anywhere on the map (except on exits or walls).\nstart_positions = [{  }]\n"
LASERS = '''
width = 6
height = 5
n_agents = 2
exits = [{ i = 4, j = 0 }, { i = 4, j = 5 }]
walls = [{ i = 2, j = 3 }, { i = 0, j = 0 }, { i = 0, j = 5 }]
voids = [{ i = 4, j = 3 }]
gems = [{ i = 3, j = 1 }]
starts = [{ row = 1 }]

[[lasers]]
direction = "East"
agent = 0
position = { i = 2, j = 0 }
laser_id = 0

[[lasers]]
direction = "S"
agent = 1
position = { i = 0, j = 4 }
laser_id = 1

[[agents]]
start_positions = [{ i = 3, j = 2 }]
'''


def _facts_oracle(text):
    w = lo.World(text)
    return dict(dims=(w.height, w.width, w.n_agents, w.n_gems, w.n_sources), walls=w.wall_pos, voids=w.void_pos, exits=w.exit_pos,
                gems=[g.pos for g in w.gems], random_starts=w.random_start_pos, laser_cells=w.laser_pos,
                sources=[(s.pos, s.agent_id, int(s.direction), s.laser_id, s.beam_len) for s in w.laser_sources],
                lasers=[(l.pos, l.laser_id, l.agent_id, int(l.direction)) for l in w.lasers])


def _facts_native(text):
    import lle_b200

    m = lle_b200.Map(text)
    return dict(dims=(m.height, m.width, m.n_agents, m.n_gems, m.n_sources), walls=m.walls, voids=m.voids, exits=m.exits,
                gems=m.gems, random_starts=m.random_starts, laser_cells=m.laser_cells,
                sources=[(s.pos, s.agent_id, int(s.direction), s.laser_id, s.beam_len) for s in m.sources()],
                lasers=[(pos, lid, colour, int(d)) for pos, lid, colour, d, _, _ in m.laser_tiles()])


@pytest.fixture(params=["oracle", "native"])
def facts(request):
    return _facts_oracle if request.param == "oracle" else _facts_native


def _errors():
    import lle_b200

    return (lo.ParsingError, lle_b200.ParsingError)


# ----------------------------------------------------------------------------- [T] src/unit_tests/test_toml_config.rs
def test_invalid_toml_field(facts):  # :5 and :19
    with pytest.raises(_errors(), match='UnknownTomlKey.*invalid_field'):
        facts('world_string = "S0 X"\ninvalid_field = 25\n')
    with pytest.raises(_errors(), match='UnknownTomlKey.*invalid_subfield'):
        facts('world_string = "S0 X"\n[[agents]]\ninvalid_subfield = 25\n')


def test_parse_toml_width_and_height_problem(facts):  # :35, :54
    with pytest.raises(_errors(), match="InconsistentWorldStringWidth.*toml_width: 10.*world_str_width: 2"):
        facts('width = 10\nworld_string = "S0 X"\n')
    with pytest.raises(_errors(), match="InconsistentWorldStringHeight.*toml_height: 10.*world_str_height: 1"):
        facts('height = 10\nworld_string = "S0 X"\n')


def test_parse_start_pos_rows_cols(facts):  # :73, :95, :116 (the maps have no exits: the config parses, the world does not)
    for starts, expected in (("[{row = 0}]", {(0, j) for j in range(10)}), ("[{col = 0}]", {(i, 0) for i in range(10)}),
                             ("[{col = 0}, {row=0}]", {(0, j) for j in range(10)} | {(i, 0) for i in range(10)})):
        f = facts(f"height = 10\nwidth = 10\nn_agents = 2\nstarts = {starts}\nexits = [{{ i = 9, j = 9 }}, {{ i = 9, j = 8 }}]\n")
        for cand in f["random_starts"]:
            assert set(cand) == expected and len(cand) == len(expected)  # duplicates are removed (19, not 20)
    with pytest.raises(_errors(), match="NotEnoughExitTiles"):
        facts("height = 10\nwidth = 10\nn_agents = 2\nstarts = [{row = 0}]\n")


def test_start_position_in_wall(facts):  # :144
    f = facts('world_string="""\n@ . .\n@ . X\n"""\n[[agents]]\nstart_positions = [{i_min=1}]\n')
    assert len(f["random_starts"][0]) == 1 and f["random_starts"] == [[(1, 1)]]


def test_ok(facts):  # :159 ; python/tests/test_world.py:680 test_world_toml
    f = facts(TEST_OK)
    assert f["dims"][:3] == (5, 10, 4)
    assert len(f["exits"]) == 6 and len(f["gems"]) == 1
    assert [len(c) for c in f["random_starts"]] == [8, 1, 2, 37]
    assert f["random_starts"][1] == [(0, 4)] and f["random_starts"][2] == [(0, 5), (3, 5)]


def test_global_start_pos(facts):  # :195
    f = facts(GLOBAL_STARTS)
    assert len(f["exits"]) == 10
    assert all(len(c) == 9 for c in f["random_starts"]) and len(f["random_starts"]) == 5


# ----------------------------------------------------------------------------- model details read off the serde derives
def test_untagged_positions_and_aliases(facts):
    # `{}` and `{ i = 3 }` are the Rect with every default (position_config.rs:5-26: untagged, unknown keys ignored)
    f = facts("width = 3\nheight = 2\nexits = [{ i = 0, j = 0 }]\n[[agents]]\nstarts = [{ i = 1 }]\n")
    assert f["random_starts"] == [[(0, 1), (0, 2), (1, 0), (1, 1), (1, 2)]]
    with pytest.raises(_errors(), match="PositionOutOfBounds"):
        facts("width = 3\nheight = 2\nexits = [{ i = 5, j = 0 }]\nn_agents = 1\nstarts = [{}]\n")
    with pytest.raises(_errors(), match="InconsistentNumberOfAgents"):
        facts('n_agents = 1\nworld_string = "S0 S1 X X"\n')
    with pytest.raises(_errors(), match="EmptyWorld"):
        facts("n_agents = 1\nexits = []\n")


def test_lasers_table(facts):  # toml_laser_config.rs:9-15, direction aliases (direction.rs:8-18)
    f = facts(LASERS)
    assert f["dims"] == (5, 6, 2, 1, 2)
    assert f["sources"] == [((2, 0), 0, 1, 0, 2), ((0, 4), 1, 2, 1, 4)]  # the east beam stops at the wall (2, 3)
    assert (2, 0) not in f["walls"]  # [[lasers]] positions are not added to the wall list (unlike v1 `L` tokens)
    # agent 1's candidates on the east beam of colour 0 are pruned; agent 0's on the south beam of colour 1 too
    assert (1, 4) not in f["random_starts"][0] and (1, 4) in f["random_starts"][1]
    assert f["random_starts"][0][-1] == (3, 2) and (3, 2) not in f["random_starts"][1]


def test_not_v2_falls_back_to_v1(facts):  # parsing/mod.rs:14-21
    assert facts("S0 . X")["dims"][:3] == (1, 3, 1)
    with pytest.raises(_errors(), match="InvalidTile"):
        facts('width = "ten"\nheight = 2\n')  # a type mismatch is "not v2"; the v1 grammar then rejects the text


def test_product_and_oracle_agree_on_toml_corpus():
    for text in (TEST_OK, GLOBAL_STARTS, ANYWHERE, LASERS):
        assert _facts_native(text) == _facts_oracle(text)


# ----------------------------------------------------------------------------- start sampling
def test_seed(api):  # [P] python/tests/test_world.py:716 test_seed: same seed, same starts
    seen = set()
    for seed in range(10):
        starts = []
        for _ in range(2):
            world = api.World(ANYWHERE)
            world.seed(seed)
            world.reset()
            starts.append(world.start_pos)
            assert world.start_pos == world.agents_positions and world.start_pos[0] in world.random_start_pos[0]
        assert starts[0] == starts[1]
        seen.add(tuple(starts[0]))
    assert len(seen) > 3  # and different seeds give different starts


def test_random_starts_are_distinct_and_allowed(api):
    world = api.World(TEST_OK)
    cands = world.random_start_pos
    for k in range(30):
        world.reset()
        pos = world.agents_positions
        assert len(set(pos)) == 4 and all(p in c for p, c in zip(pos, cands))
        assert world.start_pos == pos
    events = world.step([api.Action.STAY] * 4)
    assert events == []


def test_forced_unique_assignment(api):  # src/unit_tests/test_world.rs:553-580 / test_utils.rs:4-29: only one assignment exists
    text = 'world_string = """\n. . X\n. . X\n"""\n[[agents]]\nstarts = [{ i = 0, j = 0 }, { i = 0, j = 1 }]\n[[agents]]\nstarts = [{ i = 0, j = 0 }]\n'
    world = api.World(text)
    for _ in range(10):
        world.reset()
        assert world.agents_positions == [(0, 1), (0, 0)]


@pytest.mark.gpu
def test_random_start_maps_in_a_batch():
    """TOML maps with random starts under Philox rollouts with auto-reset: every reset of every env samples the same
    starts on the device and in the oracle (the library's own stream), and all outputs stay bit-exact."""
    from test_gpu_parity import run_pair

    run_pair([TEST_OK], None, 300, 150, seed=81)
    run_pair([GLOBAL_STARTS], None, 200, 100, seed=82, env_id_base=5000)
    run_pair([LASERS, LASERS.replace("{ i = 3, j = 2 }", "{ i = 3, j = 4 }")], [e % 2 for e in range(256)], 256, 200, seed=83,
             obs_type="partial3x3")
    ora, dev = run_pair([ANYWHERE], None, 128, 60, seed=84, auto_reset=False)
    for _ in range(3):  # explicit resets draw new starts each time
        before = np.array(ora.pos)
        ora.reset(); dev.vec.reset()
        from _parity import assert_same
        assert_same(dev, ora, dev.pull(), "explicit reset")
        assert (np.array(ora.pos) != before).any()


# ----------------------------------------------------------------------------- differential fuzz of the two TOML paths
def _random_toml(rng):
    """A random document in (and sometimes slightly outside) the v2 schema: dimensions, optional world_string, position
    configs in every form the untagged enum accepts (point, rectangle bounds, row / col), walls / voids / gems / exits / starts,
    [[agents]] with start_positions, [[lasers]] with every direction alias; now and then an unknown key, an out-of-bounds
    position, a bad direction or inconsistent dimensions, so that the error paths are compared as well."""
    H, W = rng.randint(2, 7), rng.randint(2, 8)
    n_agents = rng.randint(1, 3)

    def point():
        i, j = rng.randint(0, H - 1), rng.randint(0, W - 1)
        if rng.random() < 0.04:
            i += H  # out of bounds
        return "{ i = %d, j = %d }" % (i, j)

    def position():
        r = rng.random()
        if r < 0.5:
            return point()
        if r < 0.65:
            return "{ row = %d }" % rng.randint(0, H - 1)
        if r < 0.8:
            return "{ col = %d }" % rng.randint(0, W - 1)
        keys = []
        if rng.random() < 0.6:
            keys.append("i_min = %d" % rng.randint(0, H - 1))
        if rng.random() < 0.6:
            keys.append("i_max = %d" % rng.randint(0, H - 1))
        if rng.random() < 0.6:
            keys.append("j_min = %d" % rng.randint(0, W - 1))
        if rng.random() < 0.6:
            keys.append("j_max = %d" % rng.randint(0, W - 1))
        return "{ " + ", ".join(keys) + " }"

    def plist(n, fn=position):
        return "[" + ", ".join(fn() for _ in range(n)) + "]"

    lines = ["width = %d" % (W + (1 if rng.random() < 0.03 else 0)), "height = %d" % H]
    with_string = rng.random() < 0.4
    if not with_string or rng.random() < 0.5:
        lines.append("n_agents = %d" % n_agents)
    lines.append("exits = " + plist(rng.randint(1, 3)))
    if rng.random() < 0.6:
        lines.append("walls = " + plist(rng.randint(0, 3), point))
    if rng.random() < 0.3:
        lines.append("voids = " + plist(rng.randint(0, 2), point))
    if rng.random() < 0.5:
        lines.append("gems = " + plist(rng.randint(0, 3), point))
    if rng.random() < 0.6:
        lines.append("starts = " + plist(rng.randint(1, 2)))
    if rng.random() < 0.05:
        lines.append("colour = 3")  # deny_unknown_fields
    if with_string:
        grid = [["." for _ in range(W)] for _ in range(H)]
        for _ in range(rng.randint(0, 3)):
            grid[rng.randint(0, H - 1)][rng.randint(0, W - 1)] = rng.choice(["@", "G", "X", "V"])
        for a in range(n_agents):
            if rng.random() < 0.7:
                grid[rng.randint(0, H - 1)][rng.randint(0, W - 1)] = "S%d" % a
        lines.append('world_string = """\n' + "\n".join(" ".join(row) for row in grid) + '\n"""')
    for a in range(n_agents if rng.random() < 0.8 else rng.randint(0, 3)):
        lines.append("\n[[agents]]")
        if rng.random() < 0.6:
            lines.append("start_positions = " + plist(rng.randint(1, 2)))
    for k in range(rng.randint(0, 2)):
        lines.append("\n[[lasers]]")
        lines.append('direction = "%s"' % rng.choice(["N", "S", "E", "W", "North", "South", "East", "West", "north", "Z"][: 9 if rng.random() < 0.95 else 10]))
        lines.append("agent = %d" % rng.randint(0, n_agents - (0 if rng.random() < 0.05 else 1)))
        lines.append("position = " + point())
        if rng.random() < 0.5:
            lines.append("laser_id = %d" % k)
    return "\n".join(lines) + "\n"


def test_product_and_oracle_agree_on_random_toml_documents():
    """1,200 random documents: both paths accept the same ones, with identical worlds, and refuse the same ones.  (Documents
    the product refuses as LLE_PARSE_UNSUPPORTED - a source on an earlier beam, duplicated gems - are skipped: the reference
    leaves an inconsistent world there, see DESIGN.md.)"""
    import random
    import re

    import lle_b200

    rng = random.Random(2718)
    ok = bad = skipped = 0
    for k in range(1200):
        text = _random_toml(rng)
        native_error = oracle_error = None
        try:
            native = _facts_native(text)
        except lle_b200.ParsingError as e:
            if "nsupported" in str(e):
                skipped += 1
                continue
            native, native_error = None, re.match(r"\w+", str(e)).group(0)
        try:
            oracle = _facts_oracle(text)
        except lo.ParsingError as e:
            oracle, oracle_error = None, re.match(r"\w+", str(e)).group(0)
        except RuntimeError as e:  # the reference panics (unreachable!() in World::gems, expect() in reset ...): undefined there
            assert native is None, f"document {k}: the reference panics ({e}), the product accepted the document\n{text}"
            skipped += 1
            continue
        assert (native is None) == (oracle is None), f"document {k}: product {'refuses' if native is None else 'accepts'}, oracle does not\n{text}"
        if native is not None:
            assert native == oracle, f"document {k}\n{text}"
            ok += 1
        else:
            # the same ParseError variant (errors.rs), except where the reference would panic after its own validation
            assert native_error == oracle_error or (native_error == "PositionOutOfBounds" and oracle_error is None), (k, native_error, oracle_error, text)
            bad += 1
    assert ok > 150 and bad > 150, (ok, bad, skipped)
