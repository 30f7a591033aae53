"""The reference's own generator tests, transcribed (python/tests/test_generator.py, test_geometry_ok.py): every test that does
not need the SAT characterizer.  They run against the oracle (oracle/generator.py + the oracle engine) on CPU and against the
product (lle_b200.generator: layouts generated on the device, worlds on the device) with `-m gpu`, through one adaptor each,
so the assertions read like the reference's.  `lle_gen_geometry_valid` (a host entry point) is checked on CPU as well.
"""
import numpy as np
import pytest

from oracle import generator as og

DIR = {"N": 0, "S": 1, "E": 2, "W": 3}


# ------------------------------------------------------------------------------------------------------------- adaptors
class _OracleGen:
    def __init__(self, **kw):
        self.cfg = og.GenConfig(**kw)
        self.cfg.validate()

    def generate(self, seed, max_attempts):
        from oracle import lle_oracle as lo

        lay, _ = og.generate(self.cfg, max_attempts, seed)
        if lay is None:
            return None
        w = lo.World(lay.to_v1())
        w.reset()
        return w


class _CudaGen:
    def __init__(self, **kw):
        from lle_b200.generator import WorldGenerator

        self.g = WorldGenerator(**kw, batch=8)

    def generate(self, seed, max_attempts):
        import lle_b200

        text = self.g.generate(max_attempts, seed)
        if text is None:
            return None
        w = lle_b200.World(text)
        w.reset()
        return w


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def WorldGenerator(request):
    return _OracleGen if request.param == "oracle" else _CudaGen


def _build(gen, seed: int = 0, max_attempts: int = 500):
    world = gen.generate(seed=seed, max_attempts=max_attempts)
    assert world is not None, "Generator exhausted max_attempts without producing a world"
    return world


def _starts(world):
    return [p[0] for p in world.random_start_pos]


# ------------------------------------------------------------------------------------- python/tests/test_generator.py
def test_default_random_builds_world(WorldGenerator):  # :28
    world = _build(WorldGenerator(width=6, height=6, n_agents=2))
    assert (world.width, world.height, world.n_agents) == (6, 6, 2)


def test_single_agent(WorldGenerator):  # :36
    assert _build(WorldGenerator(width=5, height=5, n_agents=1)).n_agents == 1


def test_builder_places_requested_gems_on_free_cells(WorldGenerator):  # :42 generate(8, 8, 2).gems(5).lasers(1).walls(4).build(seed=0)
    world = _build(WorldGenerator(width=8, height=8, n_agents=2, n_gems=5, n_lasers=1, n_walls=4))
    gems = {gem.pos for gem in world.gems}
    occupied = set(_starts(world)) | set(world.exit_pos) | set(world.wall_pos) | {s.pos for s in world.laser_sources}
    assert len(gems) == 5 and not gems & occupied


def test_starts_edge_agents_on_one_edge(WorldGenerator):  # :63
    gen = WorldGenerator(width=8, height=8, n_agents=2, starts="edge", exits="random")
    for seed in range(10):
        world = _build(gen, seed=seed)
        rows, cols = [p[0] for p in _starts(world)], [p[1] for p in _starts(world)]
        assert (all(r == 0 for r in rows) or all(r == world.height - 1 for r in rows) or all(c == 0 for c in cols)
                or all(c == world.width - 1 for c in cols)), (seed, rows, cols)


def test_starts_clustered_agents_form_rectangle(WorldGenerator):  # :78
    gen = WorldGenerator(width=8, height=8, n_agents=2, starts="clustered", exits="random", cluster_shape=(1, 2))
    for seed in range(10):
        pos = _starts(_build(gen, seed=seed))
        assert max(r for r, _ in pos) - min(r for r, _ in pos) <= 1 and max(c for _, c in pos) - min(c for _, c in pos) <= 2


def test_exits_opposite_edge(WorldGenerator):  # :95
    gen = WorldGenerator(width=8, height=8, n_agents=2, starts="edge", exits="opposite")
    for seed in range(10):
        world = _build(gen, seed=seed)
        ar, ac = [p[0] for p in _starts(world)], [p[1] for p in _starts(world)]
        er, ec = [r for r, _ in world.exit_pos], [c for _, c in world.exit_pos]
        if all(c == 0 for c in ac):
            assert all(c == world.width - 1 for c in ec)
        elif all(c == world.width - 1 for c in ac):
            assert all(c == 0 for c in ec)
        elif all(r == 0 for r in ar):
            assert all(r == world.height - 1 for r in er)
        else:
            assert all(r == 0 for r in er)


def test_exits_opposite_cluster(WorldGenerator):  # :117
    gen = WorldGenerator(width=10, height=10, n_agents=2, starts="clustered", exits="opposite", cluster_shape=(1, 2))
    for seed in range(10):
        world = _build(gen, seed=seed)
        assert not set(_starts(world)) & set(world.exit_pos)


def test_exits_no_overlap_with_agents(WorldGenerator):  # :131
    for mode in ("random", "edge", "cluster"):
        world = _build(WorldGenerator(width=6, height=6, n_agents=2, exits=mode, cluster_shape=(1, 2)))
        assert not set(_starts(world)) & set(world.exit_pos), mode


def test_no_walls(WorldGenerator):  # :140
    assert _build(WorldGenerator(width=6, height=6, n_agents=2, n_walls=0)).wall_pos == []


def test_walls_individual(WorldGenerator):  # :146
    assert len(_build(WorldGenerator(width=8, height=8, n_agents=2, n_walls=5, walls_style="individual")).wall_pos) == 5


def test_walls_shapes(WorldGenerator):  # :152
    world = _build(WorldGenerator(width=8, height=8, n_agents=2, n_walls=6, walls_style="shapes"))
    assert isinstance(world.wall_pos, list) and 1 <= len(world.wall_pos) <= 6


def test_lasers_free_count(WorldGenerator):  # :164
    world = _build(WorldGenerator(width=8, height=8, n_agents=2, n_lasers=2, laser_placement="free"))
    assert len(world.laser_sources) == 2  # wall_pos holds the source cells too in a v1 map (parser_v1.rs:22-25)


def test_laser_span_int_minimum(WorldGenerator):  # :170
    world = _build(WorldGenerator(width=8, height=8, n_agents=2, n_lasers=1, laser_placement="free", laser_span=4))
    assert len(world.laser_sources) == 1 and len(world.lasers) >= 4


def test_laser_span_across(WorldGenerator):  # :178
    world = _build(WorldGenerator(width=8, height=8, n_agents=2, n_lasers=1, laser_placement="free", laser_span="across"))
    assert len(world.lasers) >= 1


def test_cross_agent_laser_crosses_all_lanes(WorldGenerator):  # :190
    world = _build(WorldGenerator(width=8, height=8, n_agents=2, starts="edge", exits="opposite", n_lasers=1, laser_placement="cross-agent"))
    assert len(world.laser_sources) == 1 and len(world.lasers) >= world.n_agents


def test_cross_agent_multiple_lasers(WorldGenerator):  # :206
    world = _build(WorldGenerator(width=10, height=10, n_agents=2, starts="edge", exits="opposite", n_lasers=2, laser_placement="cross-agent"))
    assert len(world.laser_sources) == 2


def test_cross_cluster_laser_in_corridor(WorldGenerator):  # :225
    world = _build(WorldGenerator(width=10, height=10, n_agents=2, starts="clustered", exits="opposite", n_lasers=1,
                                  laser_placement="cross-cluster", cluster_shape=(1, 2)))
    assert len(world.laser_sources) == 1


@pytest.mark.parametrize("kw, match", [
    (dict(starts="random", exits="opposite"), "opposite"),                                                          # :264
    (dict(starts="clustered", n_lasers=1, laser_placement="cross-agent"), "cross-agent"),                            # :269
    (dict(starts="edge", n_lasers=1, laser_placement="cross-cluster"), "cross-cluster"),                             # :274
    (dict(starts="clustered", exits="random", n_lasers=1, laser_placement="cross-cluster"), "cross-cluster"),        # :279
    (dict(n_lasers=1, laser_span=1), "laser_span"),                                                                  # :292
])
def test_construction_errors(kw, match):
    """The constructor checks precede any device call, so the product runs them on CPU too."""
    from lle_b200.generator import WorldGenerator as Product

    for make in (lambda **k: og.GenConfig(**k).validate(), Product):
        with pytest.raises(ValueError, match=match):
            make(width=5, height=5, n_agents=2, **kw)


def test_error_gems_exceed_cells_after_starts_and_exits():  # :297
    from lle_b200.generator import generate

    with pytest.raises(ValueError, match=r"gems must be <= grid cells minus start and exit cells \(14\)"):
        generate(width=4, height=4, n_agents=1).gems(15).lasers(0).build(max_attempts=1)


# ----------------------------------------------------------------------------------- python/tests/test_geometry_ok.py
def _layout(lasers, *, walls=(), agents=((0, 0),), exits=((4, 4),)):
    return og.Layout(5, 5, agents=list(agents), exits=list(exits), gems=[], walls=list(walls),
                     lasers=[(colour, pos, DIR[d]) for colour, pos, d in lasers])


GEOMETRY = [  # (valid, lasers, walls, exits) — test_geometry_ok.py:24-126
    (True, [], (), None),
    (True, [(0, (0, 2), "S")], (), None),
    (True, [(0, (2, 0), "E")], (), None),
    (False, [(0, (0, 2), "N")], (), None),
    (False, [(0, (4, 2), "S")], (), None),
    (False, [(0, (2, 0), "W")], (), None),
    (False, [(0, (2, 4), "E")], (), None),
    (False, [(0, (2, 0), "E")], [(2, 2)], None),
    (False, [(0, (2, 0), "E")], [(2, 1)], None),
    (False, [(0, (0, 3), "E")], (), None),
    (False, [(0, (2, 0), "E"), (1, (2, 3), "E")], (), None),
    (False, [(0, (0, 2), "S")], (), [(2, 2)]),
    (False, [(0, (0, 2), "S")], [(1, 2)], [(2, 2)]),
    (False, [(0, (1, 0), "E"), (1, (0, 2), "N")], (), None),
    (True, [(0, (0, 1), "S"), (1, (0, 3), "S")], (), None),
]


@pytest.mark.parametrize("valid, lasers, walls, exits", GEOMETRY)
def test_geometry_valid(valid, lasers, walls, exits):
    from lle_b200.generator import is_geometry_valid

    lay = _layout(lasers, walls=walls, exits=exits or ((4, 4),))
    assert og.geometry_valid(lay) is valid
    assert is_geometry_valid(np.frombuffer(lay.cell_codes(), dtype=np.uint8), 5, 5) is valid  # host entry point of the library
