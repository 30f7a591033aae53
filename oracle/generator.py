"""ORACLE (test infrastructure, never imported by lle_b200): CPU restatement of the reference's layout generator.

Restates python/lle/generator/{generator,placements,geometry,candidates,world_builder}.py of yamoling/lle v2.11.4: one
*attempt* = `WorldGenerator._try_generate(seed)` with `constraint=None` (generator.py:243-254): seed a `random.Random`,
place agents -> exits -> lasers -> walls -> gems (generator.py:188-228), check the beam geometry (candidates.py:27-41) and
serialise the grid as a v1 map (world_builder.py:83-88).  Every function cites the lines it follows.

The random stream is CPython's `random.Random` (MT19937 + `_randbelow_with_getrandbits`, `sample`, `shuffle`, `choice`,
`randint`, `choices`; Lib/random.py of CPython 3.12, the interpreter of this image) - the same generator object the
reference uses, so an attempt is a pure function of (configuration, seed).  This file uses the interpreter's own
`random.Random`; the product (lle_b200/csrc/gen_core.cuh) re-implements that stream on the device.

Pinned by tests/golden/generator_vectors.json, produced by tests/golden/make_generator_vectors.py from the REFERENCE's
placement code imported unmodified in the build container.

One contract of our own: `placements.cluster_shape` (placements.py:44-60) draws the cluster's orientation from Python's
*global* unseeded generator (`random.choice`, not `rng`), anew in each of its three callers, so the reference's clustered
modes are not a function of the seed.  Here (and in the product) the shape is an explicit configuration value used by all
three callers; the golden vectors are generated with `cluster_shape` patched to return that value.

The "needs a blocker" labels (`analyse`) are a BFS heuristic of this repository, NOT the reference's SAT-based
`Cooperative` / `Independent` predicates (world_filter.py, needs pysat): see `analyse`.
"""
from __future__ import annotations

import random
from dataclasses import dataclass

# ALL_DIRS order of placements.py:30 and the deltas of src/core/tiles/direction.rs:20-27
DIRS = ((-1, 0), (1, 0), (0, 1), (0, -1))  # N, S, E, W
DIR_LETTERS = "NSEW"
EDGES = ("left", "right", "top", "bottom")  # placements.py:87
OPPOSITE = {"left": "right", "right": "left", "top": "bottom", "bottom": "top"}  # placements.py:23-28

# geometry.py:47-63 (weight, offsets)
WALL_SHAPES = (
    (4, ((0, 0), (0, 1))), (4, ((0, 0), (1, 0))), (1, ((0, 0), (0, 1), (0, 2))), (1, ((0, 0), (1, 0), (2, 0))),
    (1, ((0, 0), (0, 1), (1, 0))), (1, ((0, 0), (0, 1), (1, 1))), (1, ((0, 0), (1, 0), (1, 1))), (1, ((0, 1), (1, 0), (1, 1))),
    (2, ((0, 0), (0, 1), (1, 0), (1, 1))),
)


class Retry(Exception):
    """placements.py:19 LayoutRetry"""


@dataclass
class GenConfig:
    """The constructor arguments of WorldGenerator (generator.py:97-186) after its validation."""
    width: int
    height: int
    n_agents: int = 2
    starts: str = "random"        # random | edge | clustered
    exits: str = "random"         # random | edge | cluster | opposite
    n_lasers: int = 0
    n_gems: int = 0
    laser_placement: str = "free"  # free | cross-agent | cross-cluster
    laser_span: object = "any"    # "any" | "across" | int >= 2
    n_walls: object = "auto"
    walls_style: str = "individual"  # individual | shapes
    n_rooms_rows: int = 0
    n_rooms_cols: int = 0
    door_size: int = 1
    cluster_shape: tuple = (1, 1)  # see the module docstring

    def validate(self):
        """generator.py:116-181; returns the resolved number of random walls."""
        c = self
        if c.exits == "opposite" and c.starts not in ("edge", "clustered"):
            raise ValueError("exits='opposite' requires starts='edge' or starts='clustered', not 'random'.")
        if c.laser_placement == "cross-agent" and c.starts != "edge":
            raise ValueError("laser_placement='cross-agent' requires starts='edge'.")
        if c.laser_placement == "cross-cluster" and c.starts != "clustered":
            raise ValueError("laser_placement='cross-cluster' requires starts='clustered'.")
        if c.laser_placement == "cross-cluster" and c.exits not in ("opposite", "cluster"):
            raise ValueError("laser_placement='cross-cluster' requires exits='opposite' or exits='cluster'.")
        if isinstance(c.laser_span, int) and c.laser_span < 2:
            raise ValueError(f"laser_span must be >= 2, got {c.laser_span}.")
        if c.width < 1 or c.height < 1:
            raise ValueError("Grid width and height must be >= 1")
        area = c.width * c.height
        if c.n_agents < 1:
            raise ValueError(f"agents must be >= 1. Got {c.n_agents}")
        if c.n_lasers < 0 or c.n_lasers > c.n_agents:
            raise ValueError(f"lasers must be in [0, agents]. Got lasers={c.n_lasers}, agents={c.n_agents}.")
        if c.n_gems < 0 or c.n_gems > area - 2 * c.n_agents:
            raise ValueError(f"gems must be in [0, {area - 2 * c.n_agents}]. Got gems={c.n_gems}.")
        if c.n_rooms_rows > 0:
            return 0
        n_walls = area // 10 if c.n_walls == "auto" else c.n_walls
        if n_walls < 0:
            raise ValueError(f"num_walls must be >= 0. Got {n_walls}")
        if n_walls >= area / 2:
            raise ValueError(f"num_walls must be < size/2. Got num_walls={n_walls}, size={area}")
        if 2 * c.n_agents + n_walls + c.n_lasers + c.n_gems > area:
            raise ValueError("layout requires more unique cells than the grid has")
        return n_walls


@dataclass
class Layout:
    """candidates.py:12-25"""
    height: int
    width: int
    agents: list
    exits: list
    gems: list
    walls: list
    lasers: list  # (colour, (i, j), dir index into DIRS)

    def to_v1(self) -> str:
        """generator.py:230-241 + world_builder.py:83-88"""
        g = [["."] * self.width for _ in range(self.height)]
        for a, (i, j) in enumerate(self.agents):
            g[i][j] = f"S{a}"
        for i, j in self.exits:
            g[i][j] = "X"
        for i, j in self.gems:
            g[i][j] = "G"
        for i, j in self.walls:
            g[i][j] = "@"
        for colour, (i, j), d in self.lasers:
            g[i][j] = f"L{colour}{DIR_LETTERS[d]}"
        return "\n".join(" ".join(row) for row in g)

    def cell_codes(self) -> bytes:
        """The device library's cell encoding (include/lle_b200.h, lle_gen_*): 0 floor, 1 wall, 2 exit, 3 gem,
        16+a start of agent a, 64 + 4*colour + dir laser source."""
        g = bytearray(self.height * self.width)
        W = self.width
        for a, (i, j) in enumerate(self.agents):
            g[i * W + j] = 16 + a
        for i, j in self.exits:
            g[i * W + j] = 2
        for i, j in self.gems:
            g[i * W + j] = 3
        for i, j in self.walls:
            g[i * W + j] = 1
        for colour, (i, j), d in self.lasers:
            g[i * W + j] = 64 + 4 * colour + d
        return bytes(g)


def room_walls(cfg: GenConfig) -> list:
    """placements.py:546-647 place_room_walls (deterministic)."""
    H, W = cfg.height, cfg.width
    nr, nc, door = cfg.n_rooms_rows, cfg.n_rooms_cols, cfg.door_size

    def stripes(total, n):
        base, rem = divmod(total - (n - 1), n)
        sizes = [base + (1 if k < rem else 0) for k in range(n)]
        starts, dividers, at = [], [], 0
        for k, s in enumerate(sizes):
            starts.append(at)
            at += s
            if k < n - 1:
                dividers.append(at)
                at += 1
        return sizes, starts, dividers

    rh, rstart, drows = stripes(H, nr)
    rw, cstart, dcols = stripes(W, nc)
    wall = [[False] * W for _ in range(H)]
    for r in drows:
        for j in range(W):
            if 0 <= r < H:
                wall[r][j] = True
    for c in dcols:
        for i in range(H):
            if 0 <= c < W:
                wall[i][c] = True
    half = door // 2
    for r in drows:
        for k in range(nc):
            mid = cstart[k] + (rw[k] - 1) // 2
            for off in range(-half, door - half):
                col = mid + off
                if cstart[k] <= col < cstart[k] + rw[k] and 0 <= r < H and 0 <= col < W:
                    wall[r][col] = False
    for c in dcols:
        for k in range(nr):
            mid = rstart[k] + (rh[k] - 1) // 2
            for off in range(-half, door - half):
                row = mid + off
                if rstart[k] <= row < rstart[k] + rh[k] and 0 <= c < W and 0 <= row < H:
                    wall[row][c] = False
    return [(i, j) for i in range(H) for j in range(W) if wall[i][j]]


class _Attempt:
    def __init__(self, cfg: GenConfig, rng: random.Random):
        self.c, self.rng = cfg, rng
        self.H, self.W = cfg.height, cfg.width
        self.n_walls = cfg.validate()
        self.reserved = set()
        self.edge = None
        self.lanes = None
        self.agent_anchor = None
        self.exit_anchor = None

    # -- helpers ---------------------------------------------------------------------------------------------------
    def free_cells(self, extra=()):
        extra = set(extra)
        return [(r, c) for r in range(self.H) for c in range(self.W) if (r, c) not in self.reserved and (r, c) not in extra]

    def beam(self, pos, d, blockers=()):
        """geometry.py:24-43 beam_tiles"""
        dr, dc = DIRS[d]
        r, c = pos[0] + dr, pos[1] + dc
        out = []
        while 0 <= r < self.H and 0 <= c < self.W and (r, c) not in blockers:
            out.append((r, c))
            r, c = r + dr, c + dc
        return out

    def span_ok(self, tiles):
        """placements.py:283-297 (with an empty wall set the "across" comparison is always true)"""
        s = self.c.laser_span
        if s == "any":
            return len(tiles) >= 2
        if s == "across":
            return True
        return len(tiles) >= s

    def candidate(self, pos, d):
        """placements.py:265-280 _make_candidate"""
        if pos in self.reserved:
            return None
        tiles = self.beam(pos, d)
        if len(tiles) < 2 or any(t in self.reserved for t in tiles) or not self.span_ok(tiles):
            return None
        return pos, d, tiles

    def cluster_cells(self, ar, ac):
        ch, cw = self.c.cluster_shape
        return [(ar + dr, ac + dc) for dr in range(ch) for dc in range(cw)][: self.c.n_agents]

    # -- stages ----------------------------------------------------------------------------------------------------
    def place_agents(self, forbidden):
        """placements.py:68-125"""
        c, rng, n = self.c, self.rng, self.c.n_agents
        if c.starts == "random":
            pool = [(r, q) for r in range(self.H) for q in range(self.W) if (r, q) not in forbidden]
            if len(pool) < n:
                raise Retry
            agents = rng.sample(pool, n)
        elif c.starts == "edge":
            self.edge = rng.choice(EDGES)
            if self.edge in ("left", "right"):
                col = 0 if self.edge == "left" else self.W - 1
                valid = [r for r in range(self.H) if (r, col) not in forbidden]
            else:
                row = 0 if self.edge == "top" else self.H - 1
                valid = [q for q in range(self.W) if (row, q) not in forbidden]
            if len(valid) < n:
                raise Retry
            self.lanes = sorted(rng.sample(valid, n))
            agents = [(r, col) for r in self.lanes] if self.edge in ("left", "right") else [(row, q) for q in self.lanes]
        elif c.starts == "clustered":
            ch, cw = c.cluster_shape
            if ch > self.H or cw > self.W:
                raise Retry
            ar = rng.randint(0, self.H - ch)
            ac = rng.randint(0, self.W - cw)
            self.agent_anchor = (ar, ac)
            agents = self.cluster_cells(ar, ac)
            if any(a in forbidden for a in agents):
                raise Retry
        else:
            raise ValueError(c.starts)
        self.reserved = set(agents) | set(forbidden)
        return agents

    def exits_on_edge(self, edge, lanes):
        """placements.py:177-205"""
        n, rng = self.c.n_agents, self.rng
        if edge in ("left", "right"):
            col = 0 if edge == "left" else self.W - 1
            if lanes is None:
                if self.H < n:
                    raise Retry
                lanes = sorted(rng.sample(range(self.H), n))
            return [(r, col) for r in lanes]
        row = 0 if edge == "top" else self.H - 1
        if lanes is None:
            if self.W < n:
                raise Retry
            lanes = sorted(rng.sample(range(self.W), n))
        return [(row, q) for q in lanes]

    def place_exits(self):
        """placements.py:133-174, 208-246"""
        c, rng, n = self.c, self.rng, self.c.n_agents
        ch, cw = c.cluster_shape
        if c.exits == "random":
            free = self.free_cells()
            if len(free) < n:
                raise Retry
            exits = rng.sample(free, n)
        elif c.exits == "edge":
            exits = self.exits_on_edge(rng.choice(list(EDGES)), None)
        elif c.exits == "cluster":
            if ch > self.H or cw > self.W:
                raise Retry
            for _ in range(64):
                ar = rng.randint(0, self.H - ch)
                ac = rng.randint(0, self.W - cw)
                exits = self.cluster_cells(ar, ac)
                if not any(e in self.reserved for e in exits):
                    self.exit_anchor = (ar, ac)
                    break
            else:
                raise Retry
        elif c.exits == "opposite":
            if self.edge is not None:
                exits = self.exits_on_edge(OPPOSITE[self.edge], self.lanes)
            elif self.agent_anchor is not None:
                ar = max(0, min(self.H - ch - self.agent_anchor[0], self.H - ch))
                ac = max(0, min(self.W - cw - self.agent_anchor[1], self.W - cw))
                self.exit_anchor = (ar, ac)
                exits = self.cluster_cells(ar, ac)
            else:
                raise Retry
        else:
            raise ValueError(c.exits)
        if any(e in self.reserved for e in exits):
            raise Retry
        self.reserved |= set(exits)
        return exits

    def select(self, cands, reserve_beam):
        """placements.py:300-331 _select_lasers"""
        n = self.c.n_lasers
        self.rng.shuffle(cands)
        chosen, sources, beams = [], set(), set()
        for pos, d, tiles in cands:
            if len(chosen) >= n:
                break
            if pos in sources or pos in beams or any(s in tiles for s in sources):
                continue
            chosen.append((pos, d))
            sources.add(pos)
            beams.update(tiles)
            self.reserved.add(pos)
            if reserve_beam:
                self.reserved.update(tiles)
        if len(chosen) < n:
            raise Retry
        return chosen

    def corridor(self, slots, d_even, d_odd, fixed_is_row):
        """placements.py:463-522 _corridor_lasers"""
        span = self.c.laser_span
        extent = self.W if fixed_is_row else self.H
        min_len = 2 if span == "any" else (0 if span == "across" else span)
        out = []
        for k, slot in enumerate(slots):
            d = d_even if k % 2 == 0 else d_odd
            forward = d in (1, 2)  # SOUTH, EAST
            at = (lambda v: (slot, v)) if fixed_is_row else (lambda v: (v, slot))
            if span == "across":
                pos = at(0 if forward else extent - 1)
                if pos in self.reserved:
                    raise Retry
            else:
                rng_v = range(0, extent - min_len) if forward else range(min_len, extent)
                valid = [v for v in rng_v if at(v) not in self.reserved]
                if not valid:
                    raise Retry
                pos = at(self.rng.choice(valid))
            tiles = self.beam(pos, d)
            if not self.span_ok(tiles):
                raise Retry
            out.append((pos, d))
            self.reserved.add(pos)
            self.reserved.update(tiles)
        return out

    def place_lasers(self):
        """placements.py:334-460"""
        c, rng, n = self.c, self.rng, self.c.n_lasers
        if n == 0:
            return []
        if c.laser_placement == "free":
            cands = []
            for r in range(self.H):
                for q in range(self.W):
                    for d in range(4):
                        nr, nq = r + DIRS[d][0], q + DIRS[d][1]
                        if not (0 <= nr < self.H and 0 <= nq < self.W):
                            continue
                        cand = self.candidate((r, q), d)
                        if cand is not None:
                            cands.append(cand)
            chosen = self.select(cands, False)
        elif c.laser_placement == "cross-agent":
            lanes = set(self.lanes)
            lo, hi = min(lanes), max(lanes)
            vertical = self.edge in ("left", "right")  # lanes are rows, beams run along columns
            n_fixed, n_other = (self.H, self.W) if vertical else (self.W, self.H)
            d_before, d_after = (1, 0) if vertical else (2, 3)  # S, N | E, W
            cands = []
            for band, d in (([f for f in range(n_fixed) if f < lo], d_before), ([f for f in range(n_fixed) if f > hi], d_after)):
                for fixed in band:
                    for other in range(n_other):
                        pos = (fixed, other) if vertical else (other, fixed)
                        cand = self.candidate(pos, d)
                        if cand is not None and lanes.issubset(t[0 if vertical else 1] for t in cand[2]):
                            cands.append(cand)
            if not cands:
                raise Retry
            chosen = self.select(cands, True)
        elif c.laser_placement == "cross-cluster":
            ch, cw = c.cluster_shape
            if self.agent_anchor is None or self.exit_anchor is None:
                raise Retry
            bottom, right = self.agent_anchor[0] + ch - 1, self.agent_anchor[1] + cw - 1
            top, left = self.exit_anchor
            if top - bottom - 1 >= n:
                rows = list(range(bottom + 1, top))
                rng.shuffle(rows)
                chosen = self.corridor(sorted(rows[:n]), 2, 3, True)
            elif left - right - 1 >= n:
                cols = list(range(right + 1, left))
                rng.shuffle(cols)
                chosen = self.corridor(sorted(cols[:n]), 1, 0, False)
            else:
                raise Retry
        else:
            raise ValueError(c.laser_placement)
        colours = rng.sample(range(c.n_agents), n)  # placements.py:366-368
        return [(col, pos, d) for (pos, d), col in zip(chosen, colours)]

    def wall_shapes(self, free):
        """geometry.py:69-99 place_wall_shapes"""
        budget = self.n_walls
        left = set(free)
        anchors = list(free)
        self.rng.shuffle(anchors)
        walls = []
        weights = [w for w, _ in WALL_SHAPES]
        shapes = [s for _, s in WALL_SHAPES]
        for anchor in anchors:
            if budget <= 0:
                break
            if anchor not in left:
                continue
            cells = None
            for shape in self.rng.choices(shapes, weights=weights, k=4):
                if len(shape) > budget:
                    continue
                trial = [(anchor[0] + dr, anchor[1] + dc) for dr, dc in shape]
                if all(t in left for t in trial):
                    cells = trial
                    break
            if cells is None:
                cells = [anchor]
            left.difference_update(cells)
            walls.extend(cells)
            budget -= len(cells)
        return walls

    def run(self) -> Layout:
        """generator.py:188-228 _make_candidate_layout"""
        c = self.c
        rooms = room_walls(c) if c.n_rooms_rows > 0 else None
        agents = self.place_agents(set(rooms) if rooms else set())
        exits = self.place_exits()
        lasers = self.place_lasers()
        if rooms is not None:
            walls = rooms
        else:
            free = self.free_cells()  # placements.py:538-543
            walls = self.wall_shapes(free) if c.walls_style == "shapes" else self.rng.sample(free, min(self.n_walls, len(free)))
        free = self.free_cells(walls)  # placements.py:258-262
        if len(free) < c.n_gems:
            raise Retry
        gems = self.rng.sample(free, c.n_gems)
        lay = Layout(self.H, self.W, agents, exits, gems, walls, lasers)
        if not geometry_valid(lay):
            raise Retry
        return lay


def geometry_valid(lay: Layout) -> bool:
    """candidates.py:27-41"""
    blockers = set(lay.walls) | {pos for _, pos, _ in lay.lasers}
    lit = set()
    for _, (r, c), d in lay.lasers:
        dr, dc = DIRS[d]
        r, c = r + dr, c + dc
        if not (0 <= r < lay.height and 0 <= c < lay.width):
            return False
        n = 0
        while 0 <= r < lay.height and 0 <= c < lay.width and (r, c) not in blockers:
            lit.add((r, c))
            n += 1
            r, c = r + dr, c + dc
        if n < 2:
            return False
    return not (set(lay.exits) & lit)


def try_generate(cfg: GenConfig, seed: int):
    """generator.py:243-254 `_try_generate(seed)` with constraint=None: a Layout, or None on LayoutRetry."""
    try:
        return _Attempt(cfg, random.Random(seed)).run()
    except Retry:
        return None


def generate(cfg: GenConfig, max_attempts: int, seed: int, require: int = 0):
    """generator.py:268-284 `generate(max_attempts, seed)`: one stream across the attempts.  `require` (label bits, see
    `analyse`) stands where the reference's constraint does (`_accept_world`, which draws no random number).
    Returns (Layout or None, attempts used)."""
    rng = random.Random(seed)
    for t in range(max_attempts):
        try:
            lay = _Attempt(cfg, rng).run()
        except Retry:
            continue
        if (analyse(lay) & require) == require:
            return lay, t + 1
    return None, max_attempts


def cells_to_v1(cells, height: int, width: int) -> str:
    """The v1 text of a cell-code grid (world_builder.py:83-88)."""
    def tok(v):
        if v < 4:
            return ".@XG"[v]
        if 16 <= v < 48:
            return f"S{v - 16}"
        return f"L{(v - 64) // 4}{DIR_LETTERS[(v - 64) % 4]}"
    return "\n".join(" ".join(tok(int(cells[r * width + c])) for c in range(width)) for r in range(height))


def attempt_seeds(seed, n: int) -> list:
    """generator.py:296-301: the per-attempt seeds of `_generate_n_multi` (`rng.randrange(sys.maxsize)`)."""
    import sys

    rng = random.Random(seed)
    return [rng.randrange(sys.maxsize) for _ in range(n)]


# --------------------------------------------------------------------------------------------------------------------
# BFS labels (heuristic of this repository; NOT the reference's SAT predicates)
# --------------------------------------------------------------------------------------------------------------------
LABEL_WALKABLE = 1      # every agent reaches >= 1 exit through non-wall, non-source cells, and the agents can be matched
#                         to distinct reachable exits
LABEL_INDEPENDENT = 2   # the same when agent a additionally avoids every cell lit by a beam of another colour (static
#                         beams as drawn at reset, nobody blocking)
LABEL_NEEDS_BLOCKER = 4  # WALKABLE and not INDEPENDENT: some agent can only get home if somebody blocks a beam


def _reach(lay: Layout, start, blocked):
    H, W = lay.height, lay.width
    seen = {start}
    todo = [start]
    while todo:
        r, c = todo.pop()
        for dr, dc in DIRS:
            p = (r + dr, c + dc)
            if 0 <= p[0] < H and 0 <= p[1] < W and p not in blocked and p not in seen:
                seen.add(p)
                todo.append(p)
    return seen


def _matchable(sets) -> bool:
    """Distinct representatives (Kuhn's augmenting paths)."""
    owner = {}

    def grow(a, visited):
        for e in sets[a]:
            if e in visited:
                continue
            visited.add(e)
            if e not in owner or grow(owner[e], visited):
                owner[e] = a
                return True
        return False

    return all(grow(a, set()) for a in range(len(sets)))


def analyse(lay: Layout) -> int:
    solid = set(lay.walls) | {pos for _, pos, _ in lay.lasers}
    beams = []
    for colour, (r, c), d in lay.lasers:
        dr, dc = DIRS[d]
        r, c = r + dr, c + dc
        while 0 <= r < lay.height and 0 <= c < lay.width and (r, c) not in solid:
            beams.append((colour, (r, c)))
            r, c = r + dr, c + dc
    exits = list(lay.exits)
    labels = 0
    plain = [[k for k, e in enumerate(exits) if e in _reach(lay, a, solid)] for a in lay.agents]
    if _matchable(plain):
        labels |= LABEL_WALKABLE
        safe = []
        for a, pos in enumerate(lay.agents):
            foreign = {cell for colour, cell in beams if colour != a}
            ok = _reach(lay, pos, solid | foreign)
            safe.append([k for k, e in enumerate(exits) if e in ok])
        labels |= LABEL_INDEPENDENT if _matchable(safe) else LABEL_NEEDS_BLOCKER
    return labels
