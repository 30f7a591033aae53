"""Generates tests/golden/generated_5x5.json: 1,024 distinct 5x5 maps with 2 agents and 2 laser sources
(BASELINE.json configs[2]) by running the REFERENCE's own placement code — python/lle/generator/{geometry,placements,
candidates}.py, imported unmodified from /root/reference with a stub for the `lle` package — with the defaults of
`lle.generate(5, 5, 2).lasers(2)` (random starts and exits, "free" laser placement, span "any", walls = area // 10 placed
individually, no gems; python/lle/generator/builder.py:69-98, generator.py:163-228) and Python's `random.Random(seed)`.

NOT applied: the `cooperative()` SAT filter (python/lle/generator/world_filter.py needs pysat, absent here).  Maps the
engine rejects (a start killed by a beam -> AgentWithoutStart) are skipped, as `WorldBuilder.build` would raise.

Usage (build container only): python tests/golden/make_generated_maps.py [/root/reference]
"""
import enum
import importlib.util
import json
import os
import random
import sys
import types

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE))]


class Direction(enum.Enum):  # the part of lle.tiles.Direction the placement code uses (src/core/tiles/direction.rs:20-27)
    NORTH = "N"
    EAST = "E"
    SOUTH = "S"
    WEST = "W"

    @property
    def delta(self):
        return {"N": (-1, 0), "E": (0, 1), "S": (1, 0), "W": (0, -1)}[self.value]


lle = types.ModuleType("lle")
tiles = types.ModuleType("lle.tiles")
tiles.Direction = Direction
typesmod = types.ModuleType("lle.types")
typesmod.Position = tuple
lle.tiles, lle.types = tiles, typesmod
pkg = types.ModuleType("lle.generator")
pkg.__path__ = [os.path.join(REF, "python", "lle", "generator")]
sys.modules.update({"lle": lle, "lle.tiles": tiles, "lle.types": typesmod, "lle.generator": pkg})


def load(name):
    spec = importlib.util.spec_from_file_location(f"lle.generator.{name}", os.path.join(pkg.__path__[0], f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


geometry = load("geometry")
placements = load("placements")
candidates = load("candidates")

H = W = 5
N_AGENTS = N_LASERS = 2
N_WALLS = (H * W) // 10
N_MAPS = 1024


def one_layout(rng):
    """generator.py:188-228 (_make_candidate_layout) without rooms."""
    ctx = placements.PlacementCtx()
    agents, reserved = placements.place_agents("random", N_AGENTS, H, W, rng, ctx, forbidden=None)
    exits, reserved = placements.place_exits("random", N_AGENTS, H, W, rng, reserved, ctx)
    lasers, reserved = placements.place_lasers(N_LASERS, "free", "any", N_AGENTS, H, W, rng, reserved, ctx)
    walls = placements.place_walls(N_WALLS, "individual", reserved, H, W, rng)
    gems, _ = placements.place_gems(0, reserved | set(walls), H, W, rng)
    layout = candidates.CandidateLayout(H, W, agents=agents, exits=exits, gems=gems, walls=walls, lasers=lasers)
    if not layout.is_geometry_valid():
        raise placements.LayoutRetry()
    return layout


def to_v1(layout):
    """world_builder.py:83-88: one token per cell, rows joined by newlines."""
    grid = [["." for _ in range(W)] for _ in range(H)]
    for a, (i, j) in enumerate(layout.agents):
        grid[i][j] = f"S{a}"
    for i, j in layout.exits:
        grid[i][j] = "X"
    for i, j in layout.gems:
        grid[i][j] = "G"
    for i, j in layout.walls:
        grid[i][j] = "@"
    for owner, (i, j), d in layout.lasers:
        grid[i][j] = f"L{owner}{d.value}"
    return "\n".join(" ".join(row) for row in grid)


from oracle import lle_oracle as lo  # noqa: E402

maps, seen, seed = [], set(), 0
while len(maps) < N_MAPS:
    rng = random.Random(seed)
    seed += 1
    try:
        text = to_v1(one_layout(rng))
    except placements.LayoutRetry:
        continue
    if text in seen:
        continue
    try:
        lo.World(text)
    except lo.ParsingError:
        continue
    seen.add(text)
    maps.append(text)
with open(os.path.join(HERE, "generated_5x5.json"), "w") as f:
    json.dump({"height": H, "width": W, "n_agents": N_AGENTS, "n_lasers": N_LASERS, "n_walls": N_WALLS, "seeds_tried": seed,
               "filter": "geometry only (cooperative() SAT filter not applied)", "maps": maps}, f, indent=0)
print(len(maps), "maps from", seed, "seeds")
print(maps[0]); print(); print(maps[1])
