"""Generates tests/golden/generator_vectors.json by running the REFERENCE's generator — python/lle/generator/
{generator,placements,geometry,candidates}.py imported unmodified from /root/reference, with stubs for the parts of the
`lle` package that need the native module (`lle.world.World`, `lle.tiles.Direction`, the characterizer) — one
`WorldGenerator._make_candidate_layout()` per seed after `rng.seed(seed)`, exactly what `_try_generate(seed)` does before
it builds the world (generator.py:243-254), with `constraint=None`.

For the clustered modes `placements.cluster_shape` is patched to return the configured shape (the reference draws it from
Python's global unseeded generator; see oracle/generator.py).

Each result is the layout in the cell encoding of include/lle_b200.h (hex), or null for a LayoutRetry.

Usage (build container only): python tests/golden/make_generator_vectors.py [/root/reference]
"""
import importlib.util
import json
import os
import sys
import types

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class Direction:
    """What the generator uses of the native lle.tiles.Direction (src/core/tiles/direction.rs:20-27, src/bindings/
    pydirection.rs): `.delta`, and `.name`, which is the single letter."""

    def __init__(self, letter, delta):
        self.name, self.delta = letter, delta


Direction.NORTH = Direction("N", (-1, 0))
Direction.EAST = Direction("E", (0, 1))
Direction.SOUTH = Direction("S", (1, 0))
Direction.WEST = Direction("W", (0, -1))


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


lle = stub("lle")
lle.__path__ = [os.path.join(REF, "python", "lle")]
stub("lle.tiles", Direction=Direction)
stub("lle.types", Position=tuple)


class StubWorld:  # stands in for the native lle.world.World: keeps the v1 string WorldBuilder.build hands over
    def __init__(self, world_str):
        self.world_string = world_str

    def reset(self):
        pass


stub("lle.world", World=StubWorld)
stub("lle.characterization")
stub("lle.characterization.world_characterization", WorldCharacterizer=object)
gen_pkg = stub("lle.generator")
gen_pkg.__path__ = [os.path.join(REF, "python", "lle", "generator")]


def load(name):
    spec = importlib.util.spec_from_file_location(f"lle.generator.{name}", os.path.join(gen_pkg.__path__[0], f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


geometry = load("geometry")
placements = load("placements")
candidates = load("candidates")
world_builder = load("world_builder")
world_filter = load("world_filter")
generator = load("generator")

DIR_INDEX = {Direction.NORTH: 0, Direction.SOUTH: 1, Direction.EAST: 2, Direction.WEST: 3}

# (name, WorldGenerator kwargs, cluster shape or None)
CONFIGS = [
    ("5x5_a2_l2", dict(width=5, height=5, n_agents=2, n_lasers=2), None),  # BASELINE config 3: lle.generate(5,5,2).lasers(2)
    ("5x5_a2_l0", dict(width=5, height=5, n_agents=2), None),
    ("4x3_a1", dict(width=4, height=3, n_agents=1, n_lasers=1, n_gems=2), None),
    ("7x6_a3_l3_g4", dict(width=7, height=6, n_agents=3, n_lasers=3, n_gems=4), None),
    ("8x8_a4_l4_span4_shapes", dict(width=8, height=8, n_agents=4, n_lasers=4, laser_span=4, walls_style="shapes", n_gems=3), None),
    ("10x10_a4_across_w20", dict(width=10, height=10, n_agents=4, n_lasers=2, laser_span="across", n_walls=20, n_gems=30), None),
    ("12x9_a8_l3_shapes", dict(width=12, height=9, n_agents=8, n_lasers=3, n_walls=25, walls_style="shapes", n_gems=6), None),
    ("6x6_edge_edge", dict(width=6, height=6, n_agents=3, starts="edge", exits="edge", n_lasers=2, n_gems=2), None),
    ("7x7_edge_opposite_cross", dict(width=7, height=7, n_agents=3, starts="edge", exits="opposite", n_lasers=2,
                                     laser_placement="cross-agent", n_walls=3), None),
    ("9x6_edge_opposite_cross_across", dict(width=9, height=6, n_agents=2, starts="edge", exits="opposite", n_lasers=2,
                                            laser_placement="cross-agent", laser_span="across", walls_style="shapes"), None),
    ("8x8_cluster_cluster", dict(width=8, height=8, n_agents=4, starts="clustered", exits="cluster", n_lasers=2, n_gems=2), (2, 2)),
    ("9x9_cluster_opposite_corridor", dict(width=9, height=9, n_agents=3, starts="clustered", exits="opposite", n_lasers=2,
                                           laser_placement="cross-cluster", n_walls=4), (1, 3)),
    ("6x10_cluster_cluster_corridor_across", dict(width=6, height=10, n_agents=2, starts="clustered", exits="cluster", n_lasers=2,
                                                  laser_placement="cross-cluster", laser_span="across", n_walls=2), (2, 1)),
    ("10x6_cluster_opposite_corridor_span3", dict(width=10, height=6, n_agents=4, starts="clustered", exits="opposite", n_lasers=3,
                                                  laser_placement="cross-cluster", laser_span=3, n_walls=0, n_gems=5), (4, 1)),
    ("11x11_rooms2x2", dict(width=11, height=11, n_agents=4, n_lasers=2, n_gems=4, n_rooms_rows=2, n_rooms_cols=2, door_size=1), None),
    ("13x10_rooms2x3_door2_edge", dict(width=13, height=10, n_agents=3, starts="edge", exits="random", n_lasers=3, n_gems=2,
                                       n_rooms_rows=2, n_rooms_cols=3, door_size=2), None),
    ("16x16_a6", dict(width=16, height=16, n_agents=6, n_lasers=6, n_gems=10, walls_style="shapes"), None),
    ("32x32_a4", dict(width=32, height=32, n_agents=4, n_lasers=4, n_gems=8), None),
]
N_SEEDS = {"32x32_a4": 24, "16x16_a6": 48}
CHAIN_CASES = {"5x5_a2_l2", "7x6_a3_l3_g4", "8x8_a4_l4_span4_shapes", "6x6_edge_edge", "9x9_cluster_opposite_corridor", "10x10_a4_across_w20"}
CHAIN_ATTEMPTS = 3
BIG_SEEDS = [2**31 - 1, 2**32, 2**32 + 12345, 2**63 - 2, 0x0123456789ABCDEF, 2**40 + 7]


def cell_codes(layout) -> str:
    g = bytearray(layout.height * layout.width)
    W = layout.width
    for a, (i, j) in enumerate(layout.agents):
        g[i * W + j] = 16 + a
    for i, j in layout.exits:
        g[i * W + j] = 2
    for i, j in layout.gems:
        g[i * W + j] = 3
    for i, j in layout.walls:
        g[i * W + j] = 1
    for owner, (i, j), d in layout.lasers:
        g[i * W + j] = 64 + 4 * owner + DIR_INDEX[d]
    return bytes(g).hex()


def main():
    out = []
    original_shape = placements.cluster_shape
    for name, kwargs, shape in CONFIGS:
        placements.cluster_shape = (lambda n, s=shape: s) if shape is not None else original_shape
        gen = generator.WorldGenerator(**kwargs)
        seeds = list(range(N_SEEDS.get(name, 96))) + BIG_SEEDS
        results = []
        for s in seeds:
            gen._rng.seed(s)
            try:
                results.append(cell_codes(gen._make_candidate_layout()))
            except placements.LayoutRetry:
                results.append(None)
        ok = sum(r is not None for r in results)
        print(f"{name}: {ok}/{len(seeds)} layouts")
        out.append({"name": name, "config": kwargs, "cluster_shape": shape, "seeds": seeds, "cells": results})
    # WorldGenerator.generate(max_attempts, seed) (generator.py:268-284): one stream across the attempts
    chains = []
    for name, kwargs, shape in CONFIGS:
        if name not in CHAIN_CASES:
            continue
        placements.cluster_shape = (lambda n, s=shape: s) if shape is not None else original_shape
        gen = generator.WorldGenerator(**kwargs)
        seeds = list(range(40)) + BIG_SEEDS[:3]
        texts = []
        for s in seeds:
            w = gen.generate(max_attempts=CHAIN_ATTEMPTS, seed=s)
            texts.append(None if w is None else w.world_string)
        print(f"chain {name}: {sum(t is not None for t in texts)}/{len(seeds)} worlds")
        chains.append({"name": name, "config": kwargs, "cluster_shape": shape, "max_attempts": CHAIN_ATTEMPTS, "seeds": seeds, "texts": texts})
    placements.cluster_shape = original_shape
    # the per-attempt seeds of _generate_n_multi (generator.py:296-301)
    import random
    rng = random.Random(2024)
    multi = [rng.randrange(sys.maxsize) for _ in range(8)]
    with open(os.path.join(HERE, "generator_vectors.json"), "w") as f:
        json.dump({"reference": "yamoling/lle v2.11.4 python/lle/generator", "python": sys.version.split()[0],
                   "attempt_seeds_2024": multi, "cases": out, "chains": chains}, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
