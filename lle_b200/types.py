"""Value types and exceptions of the reference's Python surface (python/lle/world/__init__.pyi,
src/bindings/world/{pyaction,pyevent,pyworld_state}.rs, src/bindings/pyexceptions.rs)."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field

import numpy as np


class Action(enum.IntEnum):
    """src/action.rs:9-15 ; src/bindings/world/pyaction.rs."""

    NORTH = 0
    SOUTH = 1
    EAST = 2
    WEST = 3
    STAY = 4

    @property
    def delta(self) -> tuple[int, int]:
        return _DELTAS[int(self)]

    def opposite(self) -> "Action":
        return _OPPOSITE[self]

    @staticmethod
    def variants() -> list["Action"]:
        return list(Action)

    @staticmethod
    def cardinality() -> int:
        return 5

    @staticmethod
    def from_delta(di: int, dj: int) -> "Action":
        """PyAction::from_delta (src/bindings/world/pyaction.rs:68-87).  As in the reference, the pair is read as (x, y):
        (0, -1) is NORTH and (-1, 0) is WEST, unlike `.delta`, which is (row, column)."""
        table = {(0, 0): Action.STAY, (-1, 0): Action.WEST, (1, 0): Action.EAST, (0, -1): Action.NORTH, (0, 1): Action.SOUTH}
        if (di, dj) not in table:
            raise ValueError(f"Invalid delta: ({di}, {dj}). Valid deltas for actions are (-1, 0), (1, 0), (0, -1), or (0, 1).")
        return table[(di, dj)]


_DELTAS = {0: (-1, 0), 1: (1, 0), 2: (0, 1), 3: (0, -1), 4: (0, 0)}
_OPPOSITE = {Action.NORTH: Action.SOUTH, Action.SOUTH: Action.NORTH, Action.EAST: Action.WEST, Action.WEST: Action.EAST,
             Action.STAY: Action.STAY}


class Direction(enum.IntEnum):
    """src/core/tiles/direction.rs:9-18 (numbering of include/lle_b200.h)."""

    NORTH = 0
    EAST = 1
    SOUTH = 2
    WEST = 3

    @classmethod
    def _missing_(cls, value):
        """Direction("N") / ("E") / ("S") / ("W") (src/bindings/tiles/pydirection.rs; python/tests/test_direction.py:12-23)."""
        if isinstance(value, str) and value in ("N", "E", "S", "W"):
            return (cls.NORTH, cls.EAST, cls.SOUTH, cls.WEST)["NESW".index(value)]
        raise ValueError(f"Invalid direction: {value!r}")

    @property
    def delta(self):
        """(di, dj) of one step (src/core/tiles/direction.rs:20-27)."""
        return ((-1, 0), (0, 1), (1, 0), (0, -1))[int(self)]

    def opposite(self) -> "Direction":
        return Direction((int(self) + 2) % 4)


class EventType(enum.IntEnum):
    """src/bindings/world/pyevent.rs:9-17."""

    AGENT_EXIT = 0
    GEM_COLLECTED = 1
    AGENT_DIED = 2


@dataclass(frozen=True)
class WorldEvent:
    event_type: EventType
    agent_id: int


class WorldState:
    """src/bindings/world/pyworld_state.rs:53-132."""

    def __init__(self, agents_positions, gems_collected, agents_alive=None):
        self.agents_positions = [tuple(int(x) for x in p) for p in agents_positions]
        self.gems_collected = [bool(g) for g in gems_collected]
        self.agents_alive = ([True] * len(self.agents_positions) if agents_alive is None else [bool(a) for a in agents_alive])

    def as_array(self) -> np.ndarray:
        out = [float(x) for p in self.agents_positions for x in p]
        out += [1.0 if g else 0.0 for g in self.gems_collected]
        out += [1.0 if a else 0.0 for a in self.agents_alive]
        return np.array(out, dtype=np.float32)

    @staticmethod
    def from_array(array, n_agents: int, n_gems: int) -> "WorldState":
        array = list(array)
        expected = n_agents * 3 + n_gems
        if len(array) != expected:
            raise ValueError(f"The array must have a length of {expected}.")
        pos = [(int(array[2 * i]), int(array[2 * i + 1])) for i in range(n_agents)]
        gems = [array[2 * n_agents + i] == 1.0 for i in range(n_gems)]
        alive = [array[2 * n_agents + n_gems + i] == 1.0 for i in range(n_agents)]
        return WorldState(pos, gems, alive)

    def _key(self):
        return (tuple(self.agents_positions), tuple(self.gems_collected), tuple(self.agents_alive))

    def __eq__(self, other):
        return isinstance(other, WorldState) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return f"WorldState(agents_positions={self.agents_positions}, gems_collected={self.gems_collected}, agents_alive={self.agents_alive})"


@dataclass(frozen=True)
class Agent:
    num: int
    is_dead: bool
    has_arrived: bool

    @property
    def is_alive(self) -> bool:
        return not self.is_dead


@dataclass
class Gem:
    """PyGem (src/bindings/tiles/pygem.rs): a snapshot (`pos`, `is_collected`) plus, when it comes from a world, a handle on it."""

    pos: tuple
    is_collected: bool
    _world: object = field(default=None, repr=False, compare=False)
    _index: int = field(default=-1, repr=False, compare=False)

    def collect(self):
        """PyGem.collect (pygem.rs:51-65): marks the gem collected in the world; ValueError when the tile is not a top-level gem."""
        if self._world is None:
            raise ValueError("this Gem is not bound to a world")
        self._world._collect_gem(self._index, self.pos)
        self.is_collected = True

    @property
    def agent(self):
        """PyGem.agent (pygem.rs:67-76): the agent standing on the tile, if any (None under a laser tile)."""
        return None if self._world is None else self._world._tile_agent(self.pos, gem=True)


@dataclass(frozen=True)
class Laser:
    """Snapshot of one laser tile, like PyLaser (src/bindings/tiles/pylaser.rs:44-54)."""

    pos: tuple
    laser_id: int
    agent_id: int
    direction: Direction
    is_on: bool
    is_enabled: bool
    _world: object = field(default=None, repr=False, compare=False)

    @property
    def is_off(self) -> bool:
        return not self.is_on

    @property
    def is_disabled(self) -> bool:  # pylaser.rs:67-70
        return not self.is_enabled

    @property
    def agent(self):
        """PyLaser.agent (pylaser.rs:73-81): the agent standing on the tile, if any."""
        return None if self._world is None else self._world._tile_agent(self.pos)


@dataclass(frozen=True)
class LaserSource:
    pos: tuple
    agent_id: int
    direction: Direction
    is_enabled: bool
    laser_id: int
    beam_len: int


# ---- exceptions, src/bindings/pyexceptions.rs:43-183
class InvalidWorldStateError(ValueError):
    pass


class InvalidActionError(ValueError):
    pass


class ParsingError(ValueError):
    pass


class InvalidLevelError(ValueError):
    pass


# ObservationType values (python/lle/observations.py:37-60) served by the device path -> (LLE_OBS_* kind, param, flatten)
OBS_LAYERED, OBS_PARTIAL, OBS_PERSPECTIVE, OBS_STATE = 0, 1, 2, 3
_OBS_TYPES = {
    "layered": (OBS_LAYERED, 0, False), "flattened": (OBS_LAYERED, 0, True), "layered-padded": (OBS_LAYERED, None, False),
    "layered-padded-1": (OBS_LAYERED, 1, False), "layered-padded-2": (OBS_LAYERED, 2, False), "layered-padded-3": (OBS_LAYERED, 3, False),
    "partial3x3": (OBS_PARTIAL, 3, False), "partial5x5": (OBS_PARTIAL, 5, False), "partial7x7": (OBS_PARTIAL, 7, False),
    "perspective": (OBS_PERSPECTIVE, 0, False), "state": (OBS_STATE, 0, False), "normalized-state": (OBS_STATE, 1, False),
}


def obs_spec(obs_type: str, padding_size: int = 0) -> tuple[int, int, bool]:
    """ObservationType.from_str + get_observation_generator (observations.py:62-97).  "rgb-image" needs the renderer,
    which is not on the accelerated path."""
    obs_type = getattr(obs_type, "value", obs_type)  # an ObservationType member or its string
    if obs_type == "rgb-image":
        raise NotImplementedError("observation type 'rgb-image' is not on the accelerated path")
    if obs_type not in _OBS_TYPES:
        raise ValueError(f"'{obs_type}' is not a valid ObservationType")
    kind, param, flatten = _OBS_TYPES[obs_type]
    return kind, int(padding_size) if param is None else param, flatten
