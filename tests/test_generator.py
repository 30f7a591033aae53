"""Layout generator: the reference's `WorldGenerator` attempts, one per device thread.

Pinning chain: tests/golden/generator_vectors.json holds layouts produced by the REFERENCE's own placement code
(tests/golden/make_generator_vectors.py, run in the build container) for 18 configurations covering every starts / exits /
laser placement / span / wall style / rooms mode, for `_try_generate(seed)` and for `generate(max_attempts, seed)`.
  * CPU: the oracle (oracle/generator.py, Python's own `random.Random`) reproduces them; the generator core that the kernel
    runs per thread (gen_core.cuh, __host__ __device__), instantiated on the host by tests/host_shim, reproduces them with its
    own MT19937 / sample / shuffle / choices; core vs oracle on random configurations, labels included.
  * GPU (`-m gpu`): the same comparisons through the C ABI (lle_gen_run), a million-attempt batch checked by sampling and by
    layout invariants, and the generated maps stepped on the device against the oracle engine.
"""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest

from _util import GOLDEN, build_gen_host_shim
from oracle import generator as og


@pytest.fixture(scope="module")
def vectors():
    with open(os.path.join(GOLDEN, "generator_vectors.json")) as f:
        return json.load(f)


def make_cfg(case) -> og.GenConfig:
    kw = dict(case["config"])
    if case.get("cluster_shape"):
        kw["cluster_shape"] = tuple(case["cluster_shape"])
    return og.GenConfig(**kw)


def mirror_kwargs(cfg: og.GenConfig) -> dict:
    return dict(width=cfg.width, height=cfg.height, n_agents=cfg.n_agents, starts=cfg.starts, exits=cfg.exits, n_lasers=cfg.n_lasers,
                n_gems=cfg.n_gems, laser_placement=cfg.laser_placement, laser_span=cfg.laser_span, n_walls=cfg.n_walls,
                walls_style=cfg.walls_style, n_rooms_rows=cfg.n_rooms_rows, n_rooms_cols=cfg.n_rooms_cols, door_size=cfg.door_size,
                cluster_shape=cfg.cluster_shape)


def random_config(rng: random.Random) -> og.GenConfig:
    """A valid random configuration over every mode (cluster shapes included)."""
    while True:
        w, h = rng.randint(3, 14), rng.randint(3, 14)
        n = rng.randint(1, 5)
        starts = rng.choice(["random", "edge", "clustered"])
        exits = rng.choice(["random", "edge", "cluster"] + (["opposite"] if starts != "random" else []))
        placement = "free"
        if starts == "edge" and rng.random() < 0.5:
            placement = "cross-agent"
        if starts == "clustered" and exits in ("cluster", "opposite") and rng.random() < 0.5:
            placement = "cross-cluster"
        shape = rng.choice([(1, n), (n, 1)] + ([(2, 2)] if n == 4 else []) + ([(2, 3)] if n == 5 else []))
        kw = dict(width=w, height=h, n_agents=n, starts=starts, exits=exits, n_lasers=rng.randint(0, n), n_gems=rng.randint(0, 4),
                  laser_placement=placement, laser_span=rng.choice(["any", "any", "across", 2, 3, 5]),
                  n_walls=rng.choice(["auto", 0, 1, rng.randint(0, 12)]), walls_style=rng.choice(["individual", "shapes"]), cluster_shape=shape)
        if rng.random() < 0.2 and w >= 5 and h >= 5:
            kw.update(n_rooms_rows=rng.randint(1, 2), n_rooms_cols=rng.randint(1, 3), door_size=rng.randint(1, 2))
        cfg = og.GenConfig(**kw)
        try:
            cfg.validate()
        except ValueError:
            continue
        return cfg


# ---------------------------------------------------------------------------------------------------------------- oracle
def test_oracle_reproduces_the_reference_layouts(vectors):
    total = 0
    for case in vectors["cases"]:
        cfg = make_cfg(case)
        for seed, want in zip(case["seeds"], case["cells"]):
            lay = og.try_generate(cfg, seed)
            assert (lay.cell_codes().hex() if lay else None) == want, (case["name"], seed)
            total += want is not None
    assert total > 800


def test_oracle_reproduces_the_reference_chains(vectors):
    for case in vectors["chains"]:
        cfg = make_cfg(case)
        for seed, want in zip(case["seeds"], case["texts"]):
            lay, tries = og.generate(cfg, case["max_attempts"], seed)
            assert (lay.to_v1() if lay else None) == want, (case["name"], seed)
            assert 1 <= tries <= case["max_attempts"]


def test_oracle_attempt_seeds(vectors):
    assert og.attempt_seeds(2024, 8) == vectors["attempt_seeds_2024"]


def test_generated_v1_maps_parse_in_the_oracle_engine(vectors):
    from oracle import lle_oracle as lo

    case = vectors["cases"][0]
    cfg = make_cfg(case)
    n = 0
    for seed in case["seeds"]:
        lay = og.try_generate(cfg, seed)
        if lay:
            w = lo.World(lay.to_v1())
            assert (w.height, w.width, w.n_agents) == (5, 5, 2)
            n += 1
    assert n > 50


# ------------------------------------------------------------------------------------- generator core on the host (shim)
class _Opt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in "width height n_agents starts exits n_lasers n_gems laser_placement laser_span n_walls walls_shapes "
                                        "n_rooms_rows n_rooms_cols door_size cluster_h cluster_w".split()]


def _options(cfg: og.GenConfig) -> _Opt:
    o = _Opt()
    o.width, o.height, o.n_agents = cfg.width, cfg.height, cfg.n_agents
    o.starts = {"random": 0, "edge": 1, "clustered": 2}[cfg.starts]
    o.exits = {"random": 0, "edge": 1, "cluster": 2, "opposite": 3}[cfg.exits]
    o.n_lasers, o.n_gems = cfg.n_lasers, cfg.n_gems
    o.laser_placement = {"free": 0, "cross-agent": 1, "cross-cluster": 2}[cfg.laser_placement]
    o.laser_span = {"any": 0, "across": -1}.get(cfg.laser_span, cfg.laser_span)
    o.n_walls = -1 if cfg.n_walls == "auto" else cfg.n_walls
    o.walls_shapes = int(cfg.walls_style == "shapes")
    o.n_rooms_rows, o.n_rooms_cols, o.door_size = cfg.n_rooms_rows, cfg.n_rooms_cols, cfg.door_size
    o.cluster_h, o.cluster_w = cfg.cluster_shape
    return o


class HostCore:
    """gen_core.cuh compiled by g++ (tests/host_shim/gen_host.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(build_gen_host_shim())

    def run(self, cfg, seeds, max_attempts=1, require=0):
        n, hw = len(seeds), cfg.width * cfg.height
        s = np.asarray(seeds, dtype=np.uint64)
        cells, status, labels, tries = np.zeros((n, hw), np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int32)
        err = C.create_string_buffer(256)
        o = _options(cfg)
        rc = self.lib.gen_host_run(C.byref(o), s.ctypes.data_as(C.c_void_p), C.c_int64(n), max_attempts, require,
                                   cells.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), labels.ctypes.data_as(C.c_void_p),
                                   tries.ctypes.data_as(C.c_void_p), err, 256)
        assert rc == 0, (rc, err.value)
        return cells, status, labels, tries


class DeviceCore:
    """The product: lle_gen_run through lle_b200.generator.WorldGenerator."""

    def __init__(self):
        self.gens = {}

    def run(self, cfg, seeds, max_attempts=1, require=0):
        from lle_b200.generator import WorldGenerator

        key = repr(cfg)
        if key not in self.gens:
            self.gens.clear()
            self.gens[key] = WorldGenerator(**mirror_kwargs(cfg), batch=max(4096, len(seeds)))
        g = self.gens[key]
        cells, status, labels, tries = g.run(np.asarray(seeds, dtype=np.uint64), max_attempts=max_attempts, require=require)
        return (cells.reshape(len(seeds), -1).cpu().numpy(), status.cpu().numpy(), labels.cpu().numpy(), tries.cpu().numpy())


@pytest.fixture(scope="module", params=["host-core", pytest.param("device", marks=pytest.mark.gpu)])
def core(request):
    return HostCore() if request.param == "host-core" else DeviceCore()


def test_core_reproduces_the_reference_layouts(core, vectors):
    for case in vectors["cases"]:
        cfg = make_cfg(case)
        cells, status, labels, tries = core.run(cfg, case["seeds"])
        for i, want in enumerate(case["cells"]):
            got = cells[i].tobytes().hex() if status[i] else None
            assert got == want, (case["name"], case["seeds"][i])
            assert tries[i] == 1
            if not status[i]:
                assert not cells[i].any() and labels[i] == 0


def test_core_reproduces_the_reference_chains(core, vectors):
    for case in vectors["chains"]:
        cfg = make_cfg(case)
        cells, status, _, tries = core.run(cfg, case["seeds"], max_attempts=case["max_attempts"])
        for i, want in enumerate(case["texts"]):
            got = og.cells_to_v1(cells[i], cfg.height, cfg.width) if status[i] else None
            assert got == want, (case["name"], case["seeds"][i])
            assert tries[i] == og.generate(cfg, case["max_attempts"], case["seeds"][i])[1]


def test_core_matches_the_oracle_on_random_configurations(core):
    """Differential fuzz over every mode: layouts, labels, and chains with a label requirement."""
    rng = random.Random(77)
    n_layouts = 0
    for k in range(120):
        cfg = random_config(rng)
        seeds = [rng.randrange(2**63) if rng.random() < 0.3 else rng.randrange(10**6) for _ in range(24)]
        cells, status, labels, _ = core.run(cfg, seeds)
        for i, s in enumerate(seeds):
            lay = og.try_generate(cfg, s)
            assert bool(status[i]) == (lay is not None), (cfg, s)
            if lay:
                assert cells[i].tobytes() == lay.cell_codes(), (cfg, s)
                assert labels[i] == og.analyse(lay), (cfg, s)
                n_layouts += 1
        require = rng.choice([1, 3, 5])
        cells, status, labels, tries = core.run(cfg, seeds[:8], max_attempts=5, require=require)
        for i, s in enumerate(seeds[:8]):
            lay, t = og.generate(cfg, 5, s, require)
            assert bool(status[i]) == (lay is not None) and tries[i] == t, (cfg, s, require)
            if lay:
                assert cells[i].tobytes() == lay.cell_codes() and (labels[i] & require) == require
    assert n_layouts > 600


def test_core_matches_the_oracle_at_the_limits(core):
    """Largest grids, most agents / gems / walls, degenerate 1 x N grids, rooms: chains of 3 attempts against the oracle.
    (The host instantiation of this core also runs these and 400 random configurations under ASan / UBSan cleanly.)"""
    cases = [dict(width=32, height=32, n_agents=32, n_lasers=32, n_gems=100, n_walls=400, walls_style="shapes"),
             dict(width=32, height=32, n_agents=8, n_lasers=8, n_gems=900, n_walls=0),
             dict(width=32, height=1, n_agents=3, n_lasers=2, n_walls=2), dict(width=1, height=32, n_agents=2, n_lasers=1, n_walls=3),
             dict(width=32, height=32, n_agents=4, starts="edge", exits="opposite", n_lasers=4, laser_placement="cross-agent", n_walls=300),
             dict(width=32, height=32, n_agents=4, starts="clustered", exits="opposite", n_lasers=4, laser_placement="cross-cluster",
                  cluster_shape=(2, 2), n_walls=100, n_gems=50),
             dict(width=31, height=29, n_agents=5, n_lasers=3, n_rooms_rows=3, n_rooms_cols=4, door_size=2, n_gems=10)]
    for kw in cases:
        cfg = og.GenConfig(**kw)
        seeds = list(range(10))
        cells, status, labels, tries = core.run(cfg, seeds, max_attempts=3)
        for i, s in enumerate(seeds):
            lay, t = og.generate(cfg, 3, s)
            assert bool(status[i]) == (lay is not None) and tries[i] == t, (kw, s)
            if lay:
                assert cells[i].tobytes() == lay.cell_codes() and labels[i] == og.analyse(lay), (kw, s)


def test_label_heuristic_known_cases():
    """Hand-made layouts: an open room is independent; a foreign beam across the only corridor needs a blocker; a walled-in
    agent is not walkable."""
    open_room = og.Layout(3, 4, agents=[(0, 0)], exits=[(2, 3)], gems=[], walls=[], lasers=[])
    assert og.analyse(open_room) == og.LABEL_WALKABLE | og.LABEL_INDEPENDENT
    # both exits lie beyond agent 1's beam (column 2, shot southwards from (0, 2)): agent 0 cannot cross it alone
    corridor = og.Layout(3, 5, agents=[(1, 0), (2, 0)], exits=[(1, 4), (2, 4)], gems=[], walls=[], lasers=[(1, (0, 2), 1)])
    assert og.analyse(corridor) == og.LABEL_WALKABLE | og.LABEL_NEEDS_BLOCKER
    # with one exit on the near side the agents can split: agent 1 crosses its own beam, agent 0 stays left
    split = og.Layout(3, 5, agents=[(1, 0), (2, 0)], exits=[(1, 4), (2, 1)], gems=[], walls=[], lasers=[(1, (0, 2), 1)])
    assert og.analyse(split) == og.LABEL_WALKABLE | og.LABEL_INDEPENDENT
    boxed = og.Layout(3, 3, agents=[(0, 0)], exits=[(2, 2)], gems=[], walls=[(0, 1), (1, 0), (1, 1)], lasers=[])
    assert og.analyse(boxed) == 0


# -------------------------------------------------------------------------------------------- host entry points (no GPU)
def test_attempt_seeds_and_text_through_the_abi(vectors):
    from lle_b200 import generator as G

    assert G.attempt_seeds(2024, 8).tolist() == vectors["attempt_seeds_2024"]
    seeds = G.attempt_seeds(99, 5000)
    assert seeds.tolist() == og.attempt_seeds(99, 5000)
    case = vectors["cases"][3]
    cfg = make_cfg(case)
    for seed in case["seeds"][:40]:
        lay = og.try_generate(cfg, seed)
        if lay:
            cells = np.frombuffer(lay.cell_codes(), dtype=np.uint8)
            assert G.cells_to_text(cells, cfg.height, cfg.width) == lay.to_v1()


def test_constructor_validation_matches_the_reference_messages():
    """generator.py:116-181 (the messages are the reference's); checked before any device call, so this runs without a GPU."""
    from lle_b200.generator import WorldGenerator

    cases = [
        (dict(exits="opposite"), "exits='opposite' requires starts='edge' or starts='clustered', not 'random'."),
        (dict(laser_placement="cross-agent"), "laser_placement='cross-agent' requires starts='edge'."),
        (dict(laser_placement="cross-cluster"), "laser_placement='cross-cluster' requires starts='clustered'."),
        (dict(starts="clustered", laser_placement="cross-cluster"), "laser_placement='cross-cluster' requires exits='opposite' or exits='cluster'."),
        (dict(laser_span=1), "laser_span must be >= 2, got 1."),
        (dict(width=0), "Grid width must be >= 1. Got 0"),
        (dict(n_agents=0), "agents must be >= 1. Got 0"),
        (dict(n_lasers=3), "lasers must be <= agents (one laser source per colour). Got lasers=3, agents=2."),
        (dict(n_gems=22), "gems must be <= grid cells minus start and exit cells (21). Got gems=22."),
        (dict(n_walls=13), "num_walls must be < size/2. Got num_walls=13, size=25"),
        (dict(n_walls=12, n_gems=10), "layout requires 26 unique cells, but grid has only 25"),
    ]
    for extra, message in cases:
        kw = dict(width=5, height=5, n_agents=2)
        kw.update(extra)
        with pytest.raises(ValueError) as e:
            WorldGenerator(**kw)
        assert str(e.value) == message, extra
    # and the oracle's own validation agrees on which configurations are refused
    for extra, _ in cases:
        kw = dict(width=5, height=5, n_agents=2)
        kw.update(extra)
        with pytest.raises(ValueError):
            og.GenConfig(**kw).validate()


# ------------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_million_attempts_config3_shape():
    """BASELINE config 3's generator call (5x5, 2 agents, 2 lasers) at 2^20 attempts: a sample replayed by the oracle, layout
    invariants on everything, and the acceptance rate of the reference's placement rules."""
    import torch

    from lle_b200.generator import WorldGenerator

    n = 1 << 20
    g = WorldGenerator(width=5, height=5, n_agents=2, n_lasers=2, batch=n)
    cells, status, labels, tries = g.run(first_seed=0, n=n)
    torch.cuda.synchronize()
    cfg = og.GenConfig(width=5, height=5, n_agents=2, n_lasers=2)
    st, grid, lab = status.cpu().numpy(), cells.reshape(n, 25).cpu().numpy(), labels.cpu().numpy()
    rng = random.Random(5)
    for s in [0, 1, n - 1] + [rng.randrange(n) for _ in range(1500)]:
        lay = og.try_generate(cfg, s)
        assert bool(st[s]) == (lay is not None), s
        if lay:
            assert grid[s].tobytes() == lay.cell_codes() and lab[s] == og.analyse(lay), s
    ok = grid[st == 1]
    assert 0.6 < len(ok) / n < 0.8
    assert ((ok == 2).sum(1) == 2).all() and ((ok == 16).sum(1) == 1).all() and ((ok == 17).sum(1) == 1).all()
    assert ((ok >= 64).sum(1) == 2).all() and ((ok == 1).sum(1) == 2).all() and ((ok == 3).sum(1) == 0).all()
    src = np.where(ok >= 64, (ok - 64) // 4, -1)
    assert ((src == 0).sum(1) == 1).all() and ((src == 1).sum(1) == 1).all()  # colours sampled without replacement
    assert not grid[st == 0].any() and (lab[st == 0] == 0).all() and (tries.cpu().numpy() == 1).all()
    assert ((lab[st == 1] & 1) == 1).mean() > 0.5


@pytest.mark.gpu
def test_generate_n_follows_the_parallel_path_of_the_reference():
    """generator.py:296-341: attempt i is seeded with the i-th rng.randrange(sys.maxsize); accepted attempts in order."""
    from lle_b200.generator import WorldGenerator

    cfg = og.GenConfig(width=7, height=6, n_agents=3, n_lasers=3, n_gems=4)
    g = WorldGenerator(**mirror_kwargs(cfg), batch=256)
    got = list(g.generate_n(40, seed=31337))
    want = []
    for s in og.attempt_seeds(31337, 4000):
        lay = og.try_generate(cfg, s)
        if lay:
            want.append(lay.to_v1())
        if len(want) == 40:
            break
    assert got == want
    # single-stream search (n_jobs = 1) and a bounded budget that runs out
    assert g.generate(5, seed=3) == (lambda r: r[0].to_v1() if r[0] else None)(og.generate(cfg, 5, 3))
    hard = WorldGenerator(width=4, height=3, n_agents=2, n_lasers=2, laser_span=3, n_walls=1, batch=64)
    hard_cfg = og.GenConfig(width=4, height=3, n_agents=2, n_lasers=2, laser_span=3, n_walls=1)
    for seed in range(6):
        lay, _ = og.generate(hard_cfg, 2, seed)
        assert hard.generate(2, seed=seed) == (lay.to_v1() if lay else None)


@pytest.mark.gpu
def test_generated_maps_step_bit_exactly_against_the_oracle_engine():
    """Device-generated maps (distinct, `needs a blocker` label) fed to the batched engine: 96 maps x 8 envs, 60 Philox steps,
    every output compared with the oracle engine."""
    import lle_b200
    from lle_b200.generator import NEEDS_BLOCKER, WALKABLE, WorldGenerator
    from oracle import lle_oracle as lo

    g = WorldGenerator(width=5, height=5, n_agents=2, n_lasers=2, require=WALKABLE | NEEDS_BLOCKER, batch=8192)
    maps = list(g.generate_n(96, seed=12, distinct=True))
    assert len(maps) == 96 and len(set(maps)) == 96
    cfg = og.GenConfig(width=5, height=5, n_agents=2, n_lasers=2)
    mine = []
    for s in og.attempt_seeds(12, 20000):
        lay = og.try_generate(cfg, s)
        if lay and og.analyse(lay) & 5 == 5 and lay.to_v1() not in mine:
            mine.append(lay.to_v1())
            if len(mine) == 96:
                break
    assert maps == mine
    map_of_env = [m for m in range(96) for _ in range(8)]
    vec = lle_b200.VecWorld(maps, len(map_of_env), map_of_env=map_of_env, device=0, seed=3)
    ora = lo.OracleVec(maps, map_of_env, len(map_of_env), seed=3)
    for t in range(60):
        vec.step(None)
        ora.step(None)
        for name in ("obs", "state", "avail", "reward", "done", "events", "actions", "err"):
            a, b = getattr(vec, name).cpu().numpy(), np.asarray(getattr(ora, name))
            assert np.array_equal(a, b), (t, name)


@pytest.mark.gpu
def test_generated_maps_of_other_shapes_step_against_the_oracle_engine():
    """Three more generator configurations (lanes with cross-agent lasers and wall shapes; rooms with gems; clustered starts
    with corridor lasers), 160 maps each, 4 envs per map, 40 steps: the host map compiler accepts every generated layout and
    the batched engine matches the oracle engine on all of them."""
    import lle_b200
    from lle_b200.generator import WorldGenerator
    from oracle import lle_oracle as lo

    configs = [
        dict(width=8, height=7, n_agents=3, starts="edge", exits="opposite", n_lasers=2, laser_placement="cross-agent", n_walls=5,
             walls_style="shapes", n_gems=3),
        dict(width=9, height=9, n_agents=4, n_lasers=3, n_gems=4, n_rooms_rows=2, n_rooms_cols=2, door_size=1),
        dict(width=10, height=6, n_agents=2, starts="clustered", exits="opposite", n_lasers=2, laser_placement="cross-cluster",
             cluster_shape=(2, 1), n_walls=4, n_gems=2),
    ]
    for kw in configs:
        g = WorldGenerator(**kw, batch=4096)
        maps = list(g.generate_n(160, seed=21, distinct=True))
        assert len(maps) == 160
        map_of_env = [m for m in range(len(maps)) for _ in range(4)]
        vec = lle_b200.VecWorld(maps, len(map_of_env), map_of_env=map_of_env, device=0, seed=9)
        ora = lo.OracleVec(maps, map_of_env, len(map_of_env), seed=9)
        for t in range(40):
            vec.step(None)
            ora.step(None)
            for name in ("obs", "state", "avail", "reward", "done", "events", "actions", "err"):
                a, b = getattr(vec, name).cpu().numpy(), np.asarray(getattr(ora, name))
                assert np.array_equal(a, b), (kw, t, name)


@pytest.mark.gpu
def test_builder_chain():
    """lle.generate(...).lasers(...).walls(...).take(n) (builder.py) on the device, and its World terminal."""
    from lle_b200.generator import generate

    texts = list(generate(6, 6, 3).lanes().lasers(2, placement="cross-agent").walls(3, style="shapes").gems(2).take(5, seed=8, texts=True))
    cfg = og.GenConfig(width=6, height=6, n_agents=3, starts="edge", exits="opposite", n_lasers=2, laser_placement="cross-agent", n_walls=3,
                       walls_style="shapes", n_gems=2)
    want = [lay.to_v1() for lay in (og.try_generate(cfg, s) for s in og.attempt_seeds(8, 500)) if lay][:5]
    assert texts == want
    worlds = list(generate(6, 6, 3).lanes().lasers(2, placement="cross-agent").walls(3, style="shapes").gems(2).take(2, seed=8))
    assert [w.world_string for w in worlds] == want[:2] and worlds[0].n_agents == 3 and worlds[0].n_gems == 2
    world = generate(5, 5, 2).lasers(1).rooms(2).build(seed=4, max_attempts=50)
    assert world is not None and (world.height, world.width, world.n_agents) == (5, 5, 2)
    with pytest.raises(NotImplementedError):
        generate(5, 5, 2).cooperative()


@pytest.mark.gpu
def test_run_argument_checks_and_empty_batch():
    from lle_b200 import _native
    from lle_b200.generator import WorldGenerator

    g = WorldGenerator(width=5, height=5, n_agents=2, batch=8)
    with pytest.raises(ValueError, match="max_attempts must be >= 1"):  # generator.py:279-280
        g.run(first_seed=0, n=8, max_attempts=0)
    with pytest.raises(ValueError):
        g.run(first_seed=0, n=8, require=8)
    assert _native.lib().lle_gen_run(g._h, None, 0, 9, 1, 0, None) == 202  # beyond the capacity
    cells, status, labels, tries = g.run(first_seed=0, n=0)
    assert status.numel() == 0 and cells.shape == (0, 5, 5)
    cells, status, labels, tries = g.run(first_seed=0, n=8)
    assert int(status.sum()) > 0
    # lle_gen_fetch: the host-buffer read-back for hosts without a CUDA runtime returns what the device views show
    hc, hs, hl, ht = np.zeros((3, 25), np.uint8), np.zeros(3, np.uint8), np.zeros(3, np.uint8), np.zeros(3, np.int32)
    rc = _native.lib().lle_gen_fetch(g._h, 2, 3, hc.ctypes.data, hs.ctypes.data, hl.ctypes.data, ht.ctypes.data, None)
    assert rc == 0
    assert np.array_equal(hc, cells[2:5].reshape(3, 25).cpu().numpy()) and np.array_equal(hs, status[2:5].cpu().numpy())
    assert np.array_equal(hl, labels[2:5].cpu().numpy()) and np.array_equal(ht, tries[2:5].cpu().numpy())
    assert _native.lib().lle_gen_fetch(g._h, 6, 3, None, None, None, None, None) == 202  # beyond the last run


def test_generated_layouts_compile_identically_in_product_and_oracle():
    """CPU: the v1 text of generated layouts goes through the product's host map compiler and through the oracle parser with the
    same result (cells, sources, beams, starts) - the hand-over from the generator to `lle_vec_create`."""
    from test_toml_maps import _facts_native, _facts_oracle

    rng = random.Random(99)
    n = 0
    for _ in range(40):
        cfg = random_config(rng)
        for seed in range(12):
            lay = og.try_generate(cfg, seed)
            if lay is None:
                continue
            text = lay.to_v1()
            assert _facts_native(text) == _facts_oracle(text), (cfg, seed)
            n += 1
    assert n > 150
