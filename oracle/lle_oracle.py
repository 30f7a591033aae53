"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/_build/liblle_oracle.so.

Exposes the oracle behind the same Python surface as the reference's ``lle`` module
(``World``, ``WorldState``, ``Action``, ``EventType``, ``WorldEvent``, ``LLE``; reference:
python/lle/world/__init__.pyi, src/bindings/world/pyworld.rs:144-626) so that the reference's
known-answer tests can be transcribed 1:1 (tests/kat_*.py), plus ``OracleVec``: N lock-stepped
environments whose output arrays have the layout of the device buffers of ``lle_b200``.

Nothing under ``lle_b200/`` may import this module.
"""

from __future__ import annotations

import ctypes as C
import enum
import os
import subprocess
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblle_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++ only, seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "lle_oracle.hpp", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.lleo_last_error.restype = C.c_char_p
        for name in ("lleo_world_new", "lleo_env_new", "lleo_vec_new", "lleo_vec_extras", "lleo_vec_reward"):
            getattr(L, name).restype = C.c_void_p
        L.lleo_vec_rollout.restype = C.c_double
        L.lleo_vec_step_count.restype = C.c_uint64
        _lib = L
    return _lib


# ------------------------------------------------------------------ value types
class Action(enum.IntEnum):
    """src/action.rs:9-15."""

    NORTH = 0
    SOUTH = 1
    EAST = 2
    WEST = 3
    STAY = 4

    @property
    def delta(self):
        return {0: (-1, 0), 1: (1, 0), 2: (0, 1), 3: (0, -1), 4: (0, 0)}[int(self)]

    @staticmethod
    def variants():
        return list(Action)

    @staticmethod
    def from_delta(di: int, dj: int) -> "Action":
        """src/bindings/world/pyaction.rs:68-87 (the pair is (x, y) there: (0, -1) is NORTH, (-1, 0) is WEST)."""
        table = {(0, 0): Action.STAY, (-1, 0): Action.WEST, (1, 0): Action.EAST, (0, -1): Action.NORTH, (0, 1): Action.SOUTH}
        if (di, dj) not in table:
            raise ValueError(f"Invalid delta: ({di}, {dj}). Valid deltas for actions are (-1, 0), (1, 0), (0, -1), or (0, 1).")
        return table[(di, dj)]


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
        self.agents_alive = [True] * len(self.agents_positions) if agents_alive is None else [bool(a) for a in agents_alive]

    def as_array(self) -> np.ndarray:
        out = []
        for i, j in self.agents_positions:
            out += [float(i), float(j)]
        out += [1.0 if g else 0.0 for g in self.gems_collected]
        out += [1.0 if a else 0.0 for a in self.agents_alive]
        return np.array(out, dtype=np.float32)

    @staticmethod
    def from_array(array, n_agents: int, n_gems: int) -> "WorldState":
        array = list(array)
        if len(array) != n_agents * 3 + n_gems:
            raise ValueError(f"The array must have a length of {n_agents * 3 + n_gems}.")
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
        return f"WorldState({self.agents_positions}, {self.gems_collected}, {self.agents_alive})"


# ------------------------------------------------------------------ exceptions (src/bindings/pyexceptions.rs:43-183)
class InvalidWorldStateError(ValueError):
    pass


class InvalidActionError(ValueError):
    pass


class ParsingError(ValueError):
    pass


class InvalidLevelError(ValueError):
    pass


_PARSE_KINDS = range(1, 100)
_RT = dict(InvalidAction=101, InvalidNumberOfGems=102, InvalidNumberOfAgents=103, InvalidAgentPosition=104,
           OutOfWorldPosition=105, InvalidNumberOfActions=106, InvalidWorldState=107, TileNotWalkable=108, Panic=109)


class RustPanic(RuntimeError):
    """A code path on which the reference engine would panic."""


def _raise(status: int):
    msg = lib().lleo_last_error().decode()
    if status in _PARSE_KINDS:
        raise ParsingError(msg)
    if status == _RT["InvalidAction"]:
        raise InvalidActionError(msg)
    if status in (_RT["InvalidNumberOfGems"], _RT["InvalidNumberOfAgents"], _RT["InvalidAgentPosition"],
                  _RT["InvalidWorldState"]):
        raise InvalidWorldStateError(msg)
    if status == _RT["OutOfWorldPosition"] or status == 201:
        raise IndexError(msg)
    if status in (_RT["InvalidNumberOfActions"], 202):
        raise ValueError(msg)
    if status == _RT["Panic"]:
        raise RustPanic(msg)
    raise RuntimeError(f"oracle error {status}: {msg}")


def _check(status: int):
    if status != 0:
        _raise(status)


# ------------------------------------------------------------------ map text (parsing/mod.rs:14-21: TOML v2 first, then v1)
def v1_config(world_str: str) -> dict:
    """parse_v1 (parser_v1.rs:132-175) as a dict of position lists; raises ParsingError like the reference."""
    buf = C.create_string_buffer(1 << 20)
    _check(lib().lleo_v1_config_text(world_str.encode(), buf, len(buf)))
    tok = buf.value.decode().split()
    k = 0

    def take(name):
        nonlocal k
        assert tok[k] == name, (tok[k], name)
        n = int(tok[k + 1])
        out = [(int(tok[k + 2 + 2 * q]), int(tok[k + 3 + 2 * q])) for q in range(n)]
        k += 2 + 2 * n
        return out

    assert tok[0] == "%LLE-CONFIG" and tok[1] == "size"
    cfg = dict(height=int(tok[2]), width=int(tok[3]))
    k = 4
    for name in ("gems", "voids", "exits", "walls"):
        cfg[name] = take(name)
    n_agents = int(tok[k + 1])
    k += 2
    cfg["starts"] = [take("starts") for _ in range(n_agents)]
    n_lasers = int(tok[k + 1])
    k += 2
    cfg["lasers"] = []
    for _ in range(n_lasers):
        cfg["lasers"].append(tuple(int(x) for x in tok[k + 1:k + 6]))
        k += 6
    return cfg


def prepare_map_text(text: str) -> bytes:
    """What the C++ oracle is given for a map: v1 text as is, a TOML v2 document as the config text of its WorldConfig."""
    from . import toml_config

    try:
        cfg = toml_config.to_config_text(text)
    except toml_config.TomlMapError as e:
        raise ParsingError(str(e)) from None
    return (cfg if cfg is not None else text).encode()


# ------------------------------------------------------------------ levels
def level_text(n: int) -> str:
    """The six built-in maps, committed as fixtures under tests/golden/levels/ (copied verbatim from
    the reference's resources/levels/lvl1..6 by tests/golden/make_fixtures.py)."""
    if not 1 <= n <= 6:
        raise InvalidLevelError(f"InvalidLevel {{ asked: {n}, min: 1, max: 6 }}")
    path = os.path.join(os.path.dirname(_HERE), "tests", "golden", "levels", f"lvl{n}")
    with open(path) as f:
        return f.read()


# ------------------------------------------------------------------ World
@dataclass(frozen=True)
class Agent:
    num: int
    is_dead: bool
    has_arrived: bool

    @property
    def is_alive(self):
        return not self.is_dead


@dataclass
class Gem:
    """PyGem (src/bindings/tiles/pygem.rs)."""
    pos: tuple
    is_collected: bool
    _world: object = field(default=None, repr=False, compare=False)

    def collect(self):  # pygem.rs:51-65
        rc = lib().lleo_world_gem_collect(self._world._h, C.c_long(self.pos[0]), C.c_long(self.pos[1]))
        if rc:
            raise ValueError(f"Tile at {self.pos} is not a gem")
        self.is_collected = True

    @property
    def agent(self):  # pygem.rs:67-76
        if self.pos in {l.pos for l in self._world.lasers}:
            return None
        a = lib().lleo_world_tile_agent(self._world._h, C.c_long(self.pos[0]), C.c_long(self.pos[1]))
        return None if a < 0 else a


class Direction(enum.IntEnum):
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


@dataclass(frozen=True)
class Laser:
    pos: tuple
    laser_id: int
    agent_id: int
    direction: Direction
    is_on: bool
    is_enabled: bool
    _world: object = field(default=None, repr=False, compare=False)

    @property
    def is_off(self):
        return not self.is_on

    @property
    def is_disabled(self):
        return not self.is_enabled

    @property
    def agent(self):  # pylaser.rs:73-81
        a = lib().lleo_world_tile_agent(self._world._h, C.c_long(self.pos[0]), C.c_long(self.pos[1]))
        return None if a < 0 else a


class LaserSource:
    """PyLaserSource (src/bindings/tiles/pylaser_source.rs): snapshot + handle on the world."""

    def __init__(self, world: "World", idx: int, rec):
        self._world, self._idx = world, idx
        self.pos = (int(rec[0]), int(rec[1]))
        self._agent_id = int(rec[2])
        self.direction = Direction(int(rec[3]))
        self._enabled = bool(rec[4])
        self.laser_id = int(rec[5])
        self.beam_len = int(rec[6])

    def _set_status(self, enabled: bool):  # pylaser_source.rs:55-74
        if self._enabled == bool(enabled):
            return
        lib().lleo_world_source_set_enabled(self._world._h, self._idx, int(bool(enabled)))
        self._enabled = bool(enabled)

    is_enabled = property(lambda self: self._enabled, lambda self, v: self._set_status(bool(v)))
    is_disabled = property(lambda self: not self._enabled, lambda self, v: self._set_status(not v))

    def disable(self):
        self._set_status(False)

    def enable(self):
        self._set_status(True)

    def __eq__(self, other):  # agent id, direction, laser id and position (pylaser_source.rs:144-152)
        return (isinstance(other, LaserSource) and self.agent_id == other.agent_id and self.direction == other.direction
                and self.laser_id == other.laser_id and self.pos == other.pos)

    __hash__ = None

    @property
    def agent_id(self) -> int:
        return self._agent_id

    @agent_id.setter
    def agent_id(self, colour: int):
        # src/bindings/tiles/pylaser_source.rs:107-142 (validation of the python setter)
        if colour < 0:
            raise OverflowError("can't convert negative int to unsigned")
        w = self._world
        if colour >= w.n_agents:
            raise ValueError("Agent ID is greater than the number of agents")
        lib().lleo_world_source_set_agent_id(w._h, self._idx, colour)
        cells = {l.pos for l in w.lasers if l.laser_id == self.laser_id}
        for start_agent, starts in enumerate(w.random_start_pos):
            if start_agent != colour and cells & set(starts):
                raise ValueError(f"Laser source cannot be changed to agent ID {colour}")
        self._agent_id = colour

    def set_colour(self, colour: int):
        self.agent_id = colour


class World:
    """Single oracle world behind the reference's ``lle.World`` surface."""

    def __init__(self, map_str: str):
        st = C.c_int(0)
        self._h = C.c_void_p(lib().lleo_world_new(prepare_map_text(map_str), C.byref(st)))
        _check(st.value)
        self.map_str = map_str
        d = (C.c_int * 8)()
        lib().lleo_world_dims(self._h, d)
        self.height, self.width, self.n_agents, self.n_gems, self.n_sources = d[0], d[1], d[2], d[3], d[4]
        self._random_starts = [[p] for p in self._positions(4)]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().lleo_world_free(h)
            self._h = None

    @staticmethod
    def level(n: int) -> "World":
        return World(level_text(n))

    @staticmethod
    def from_file(name: str) -> "World":
        low = name.lower()
        for prefix in ("lvl", "level"):
            if low.startswith(prefix) and low[len(prefix):].isdigit():
                return World.level(int(low[len(prefix):]))
        with open(name) as f:
            return World(f.read())

    def _positions(self, kind: int):
        buf = (C.c_int * (2 * self.height * self.width + 2 * 64))()
        n = lib().lleo_world_positions(self._h, kind, buf, len(buf) // 2)
        return [(buf[2 * k], buf[2 * k + 1]) for k in range(n)]

    # --- core API
    def reset(self):
        _check(lib().lleo_world_reset(self._h))

    def step(self, actions) -> list[WorldEvent]:
        if isinstance(actions, Action):
            actions = [actions]
        elif not isinstance(actions, (list, tuple)) or not all(isinstance(a, Action) for a in actions):
            raise TypeError("Action must be of type Action or list[Action]")
        acts = (C.c_uint8 * max(1, len(actions)))(*[int(a) for a in actions])
        ev = (C.c_int * (4 * self.n_agents * (self.n_agents + 2)))()
        ps = (C.c_int * (2 * self.n_agents * (self.n_agents + 2)))()
        n = C.c_int(0)
        _check(lib().lleo_world_step(self._h, acts, len(actions), ev, ps, C.byref(n)))
        self.last_event_passes = [ps[k] for k in range(n.value)]
        return [WorldEvent(EventType(ev[2 * k]), ev[2 * k + 1]) for k in range(n.value)]

    def available_actions(self) -> list[list[Action]]:
        mask = (C.c_uint8 * (5 * self.n_agents))()
        order = (C.c_int8 * (5 * self.n_agents))()
        lib().lleo_world_available(self._h, mask, order)
        return [[Action(order[a * 5 + k]) for k in range(5) if order[a * 5 + k] >= 0] for a in range(self.n_agents)]

    def set_agents_positions(self, agents_positions) -> list[WorldEvent]:
        """PyWorld::set_agents_positions (pyworld.rs:252-263): the current state with new positions, through set_state."""
        state = self.get_state()
        state.agents_positions = [tuple(int(x) for x in p) for p in agents_positions]
        return self.set_state(state)

    def set_agent_position(self, agent_id: int, position) -> list[WorldEvent]:
        """PyWorld::set_agent_position (pyworld.rs:282-299)."""
        if agent_id < 0:
            raise OverflowError("can't convert negative int to unsigned")
        if agent_id >= self.n_agents:
            raise ValueError(f"Agent id {agent_id} is out of bounds")
        state = self.get_state()
        state.agents_positions[agent_id] = tuple(int(x) for x in position)
        return self.set_state(state)

    def gem_at(self, position) -> Gem:
        """PyWorld::gem_at (pyworld.rs:315-327): only a top-level Gem tile qualifies; a gem wrapped by a laser tile is a
        `Tile::Laser` there and raises like any other tile."""
        i, j = (int(x) for x in position)
        if not (0 <= i < self.height and 0 <= j < self.width):
            raise IndexError("Position out of bounds")
        on_beam = {l.pos for l in self.lasers}
        for gem in self.gems:
            if gem.pos == (i, j) and (i, j) not in on_beam:
                return gem
        raise ValueError(f"Tile at position {(i, j)} is not a gem")

    def save(self, filename: str) -> None:
        """PyWorld::save (pyworld.rs:183-189)."""
        try:
            with open(filename, "w") as f:
                f.write(self.world_string)
        except OSError as e:
            raise ValueError(f"Could not write to file: {filename}: {e}") from e

    def available_joint_actions(self) -> list[list[Action]]:
        """world.rs:257-263"""
        import itertools

        return [list(joint) for joint in itertools.product(*self.available_actions())]

    # world.rs:645-652 `impl Clone for World`; pyworld.rs:557-626 __deepcopy__ / __getstate__ / __setstate__
    def _config(self):
        return (self.map_str, [(s.agent_id, s.is_enabled) for s in self.laser_sources], self.get_state())

    @staticmethod
    def _from_config(text, sources, state):
        w = World(text)
        for src, (agent_id, enabled) in zip(w.laser_sources, sources):
            if src.agent_id != agent_id:
                src.agent_id = agent_id
            if src.is_enabled != enabled:
                src.is_enabled = enabled
        w.set_state(state)
        return w

    def __deepcopy__(self, _memo) -> "World":
        return World._from_config(*self._config())

    def __reduce__(self):
        return (World._from_config, self._config())

    def available_mask(self) -> np.ndarray:
        mask = (C.c_uint8 * (5 * self.n_agents))()
        lib().lleo_world_available(self._h, mask, None)
        return np.frombuffer(mask, dtype=np.uint8).reshape(self.n_agents, 5).astype(bool)

    def get_state(self) -> WorldState:
        pos = (C.c_int * (2 * self.n_agents))()
        gems = (C.c_uint8 * max(1, self.n_gems))()
        alive = (C.c_uint8 * self.n_agents)()
        lib().lleo_world_get_state(self._h, pos, gems, alive)
        return WorldState([(pos[2 * a], pos[2 * a + 1]) for a in range(self.n_agents)],
                          [bool(gems[g]) for g in range(self.n_gems)], [bool(alive[a]) for a in range(self.n_agents)])

    def set_state(self, state: WorldState) -> list[WorldEvent]:
        na, ng = len(state.agents_positions), len(state.gems_collected)
        pos = (C.c_long * max(1, 2 * na))(*[int(x) for p in state.agents_positions for x in p])
        gems = (C.c_uint8 * max(1, ng))(*[int(g) for g in state.gems_collected])
        alive = (C.c_uint8 * max(1, na, len(state.agents_alive)))(*[int(a) for a in state.agents_alive])
        ev = (C.c_int * (4 * max(1, self.n_agents)))()
        n = C.c_int(0)
        _check(lib().lleo_world_set_state(self._h, pos, na, gems, ng, alive, ev, C.byref(n)))
        return [WorldEvent(EventType(ev[2 * k]), ev[2 * k + 1]) for k in range(n.value)]

    # --- getters
    @property
    def agents_positions(self):
        return self._positions(6)

    @property
    def agents(self):
        dead = (C.c_uint8 * self.n_agents)()
        arr = (C.c_uint8 * self.n_agents)()
        lib().lleo_world_agents(self._h, dead, arr)
        return [Agent(a, bool(dead[a]), bool(arr[a])) for a in range(self.n_agents)]

    @property
    def gems(self):
        flags = (C.c_uint8 * max(1, self.n_gems))()
        if lib().lleo_world_gem_flags(self._h, flags):
            raise RuntimeError("unreachable: a gem position holds no gem tile (World::gems panics there, world.rs:129-139)")
        return [Gem(p, bool(flags[g]), self) for g, p in enumerate(self._positions(3))]

    @property
    def gems_collected(self) -> int:
        return lib().lleo_world_n_gems_collected(self._h)

    @property
    def lasers(self):
        cap = 2 * self.height * self.width + 8
        buf = (C.c_int * (7 * cap))()
        n = lib().lleo_world_lasers(self._h, buf, cap)
        if n < 0:
            raise RuntimeError("unreachable: a laser position holds no laser tile (World::lasers panics there, world.rs:159-172)")
        return [Laser((buf[7 * k], buf[7 * k + 1]), buf[7 * k + 2], buf[7 * k + 3], Direction(buf[7 * k + 4]),
                      bool(buf[7 * k + 5]), bool(buf[7 * k + 6]), self) for k in range(n)]

    @property
    def laser_sources(self):
        buf = (C.c_int * (7 * max(1, self.n_sources)))()
        n = lib().lleo_world_sources(self._h, buf, self.n_sources)
        return [LaserSource(self, k, buf[7 * k:7 * k + 7]) for k in range(n)]

    def source_at(self, pos):
        for s in self.laser_sources:
            if s.pos == tuple(pos):
                return s
        raise ValueError(f"No laser source at {pos}")

    world_string = property(lambda self: self.map_str)
    wall_pos = property(lambda self: self._positions(0))
    void_pos = property(lambda self: self._positions(1))
    def _set_exit_pos(self, exits):  # PyWorld.exit_pos setter (pyworld.rs:202-210) -> World::set_exit_positions (world.rs:195-234)
        flat = (C.c_long * max(1, 2 * len(exits)))(*[int(x) for p in exits for x in p])
        _check(lib().lleo_world_set_exits(self._h, flat, len(exits)))

    exit_pos = property(lambda self: self._positions(2), _set_exit_pos)
    start_pos = property(lambda self: self._positions(4))
    laser_pos = property(lambda self: self._positions(5))

    @property
    def random_start_pos(self):
        """World::possible_starts (world.rs:297-303)."""
        buf = (C.c_long * 200000)()
        n = lib().lleo_world_random_starts(self._h, buf, len(buf))
        out, k = [], 0
        while k < n:
            cnt = buf[k]
            out.append([(int(buf[k + 1 + 2 * q]), int(buf[k + 2 + 2 * q])) for q in range(cnt)])
            k += 1 + 2 * cnt
        return out

    def seed(self, seed_value: int):
        """World::seed (world.rs:92-96); the start sampler's stream is the device library's contract (lle_oracle.hpp)."""
        lib().lleo_world_seed(self._h, C.c_uint64(int(seed_value) & 0xFFFFFFFFFFFFFFFF))

    @property
    def n_laser_colours(self):
        return len({s.agent_id for s in self.laser_sources})

    # --- white-box helpers for parity checks
    def beam_bits(self, idx: int) -> list[bool]:
        n = self.laser_sources[idx].beam_len
        buf = (C.c_uint8 * max(1, n))()
        lib().lleo_world_beam_bits(self._h, idx, buf)
        return [bool(buf[k]) for k in range(n)]

    def occupied(self) -> np.ndarray:
        buf = (C.c_uint8 * (self.height * self.width))()
        lib().lleo_world_occupied(self._h, buf)
        return np.frombuffer(buf, dtype=np.uint8).reshape(self.height, self.width).astype(bool)

    def observe_layered(self) -> np.ndarray:
        """(A, C, H, W) float32, python/lle/observations.py:254-266."""
        c = 2 * self.n_agents + 4
        out = np.zeros((c, self.height, self.width), dtype=np.float32)
        _check(lib().lleo_world_observe_layered(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return np.tile(out, (self.n_agents, 1, 1, 1))

    def observe(self, obs_type: str = "layered", padding_size: int = 0) -> np.ndarray:
        """`ObservationType(obs_type).get_observation_generator(world, padding_size).observe()` (observations.py:66-97)
        with a generator built from the live world: (n_agents[+padding], *shape) float32."""
        kind, param, flatten = obs_spec(obs_type, padding_size)
        out, shape6 = _observe_into(lambda buf, cap, s6: lib().lleo_world_observe(self._h, kind, param, buf, cap, s6))
        return _finish_obs(out, shape6, flatten, self.n_agents)

    def state_array(self) -> np.ndarray:
        out = np.zeros(3 * self.n_agents + self.n_gems, dtype=np.float32)
        lib().lleo_world_state_array(self._h, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out


def obs_spec(obs_type: str, padding_size: int = 0) -> tuple[int, int, bool]:
    """ObservationType value (observations.py:37-60) -> (kind, param, flatten) of ObsGen."""
    table = {"layered": (0, 0, False), "flattened": (0, 0, True), "layered-padded": (0, int(padding_size), False),
             "layered-padded-1": (0, 1, False), "layered-padded-2": (0, 2, False), "layered-padded-3": (0, 3, False),
             "partial3x3": (1, 3, False), "partial5x5": (1, 5, False), "partial7x7": (1, 7, False),
             "perspective": (2, 0, False), "state": (3, 0, False), "normalized-state": (3, 1, False)}
    if obs_type not in table:
        raise ValueError(f"'{obs_type}' is not a valid ObservationType")
    return table[obs_type]


def _observe_into(call):
    s6 = (C.c_long * 6)()
    _check(call(None, 0, s6))  # shape query
    n = int(np.prod([s6[k] for k in range(1, 1 + s6[0])]))
    out = np.zeros(n, dtype=np.float32)
    _check(call(out.ctypes.data_as(C.POINTER(C.c_float)), n, s6))
    return out, list(s6)


def _finish_obs(flat: np.ndarray, s6, flatten: bool, n_agents: int | None = None) -> np.ndarray:
    """One env's block -> the array the reference returns: np.tile over agents where the generator tiles."""
    block = flat.reshape([s6[k] for k in range(1, 1 + s6[0])])
    if s6[0] == 1:  # state: np.tile(state, (n_agents, 1))
        return np.tile(block, (n_agents, 1))
    if s6[5]:
        block = np.tile(block, (s6[5], 1, 1, 1))
    return block.reshape(block.shape[0], -1) if flatten else block


# ------------------------------------------------------------------ LLE (python/lle/env/env.py)
@dataclass
class Step:
    obs: np.ndarray
    available_actions: np.ndarray
    state: np.ndarray
    reward: np.ndarray
    done: bool
    events: list
    info: dict = field(default_factory=dict)


def _src_list(sources):
    """None -> all sources (n = -1); otherwise a list of indices into World::sources() order."""
    if sources is None:
        return -1, None
    arr = (C.c_int * max(1, len(sources)))(*[int(x) for x in sources])
    return len(sources), arr


class LLE:
    def __init__(self, map_str: str, multi_objective: bool = False, walkable_lasers: bool = True, extras=None, pbrs=None,
                 obs_type: str = "layered", padding_size: int = 0, randomize_lasers: bool = False, state_type: str = "state"):
        """extras: None | "laser_subgoal" | list of source indices.  pbrs: None | dict(gamma=0.99, reward_value=0.5,
        lasers_to_reward=None (all) | list of source indices, with_extras=True) — Builder.pbrs (builder.py:77-110)."""
        self._ctor = dict(map_str=map_str, multi_objective=multi_objective, walkable_lasers=walkable_lasers, extras=extras, pbrs=pbrs,
                          obs_type=obs_type, padding_size=padding_size, randomize_lasers=randomize_lasers, state_type=state_type)
        st = C.c_int(0)
        self._h = C.c_void_p(lib().lleo_env_new(prepare_map_text(map_str), int(multi_objective), int(walkable_lasers), C.byref(st)))
        _check(st.value)
        extras_src = None if extras in (None, "laser_subgoal") else list(extras)
        want_extras = extras is not None
        if pbrs is not None:
            n, arr = _src_list(pbrs.get("lasers_to_reward"))
            if pbrs.get("with_extras", True) and not want_extras:
                want_extras, extras_src = True, pbrs.get("lasers_to_reward")
            _check(lib().lleo_env_enable_pbrs(self._h, C.c_double(pbrs.get("gamma", 0.99)), C.c_double(pbrs.get("reward_value", 0.5)), n, arr))
        if want_extras:
            n, arr = _src_list(extras_src)
            _check(lib().lleo_env_enable_extras(self._h, n, arr))
        d = (C.c_int * 6)()
        lib().lleo_env_dims(self._h, d)
        self.height, self.width, self.n_agents, self.n_gems, self.n_channels, self.reward_dim = list(d)
        self.reward_dim = lib().lleo_env_reward_dim(self._h)
        self.extras_dim = lib().lleo_env_extras_dim(self._h)
        lib().lleo_env_set_randomize_lasers(self._h, int(bool(randomize_lasers)))
        kind, param, self._flatten = obs_spec(obs_type, padding_size)
        self._s6 = (C.c_long * 6)()
        _check(lib().lleo_env_set_obs(self._h, kind, param, self._s6))
        lib().lleo_env_obs_floats.restype = C.c_long
        lib().lleo_env_state_obs_floats.restype = C.c_long
        self.state_type = state_type
        self._state_s6 = None
        if state_type != "state":  # Builder.state_type (builder.py:51-58): any ObservationType as the state
            kind, param, self._state_flatten = obs_spec(state_type, 0)  # built with padding_size = 0 (env.py:86)
            self._state_s6 = (C.c_long * 6)()
            _check(lib().lleo_env_set_state_type(self._h, kind, param, self._state_s6))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().lleo_env_free(h)
            self._h = None

    @staticmethod
    def level(n: int, **kw) -> "LLE":
        return LLE(level_text(n), **kw)

    def __deepcopy__(self, _memo) -> "LLE":
        clone = LLE(**self._ctor)
        clone.set_state(WorldState.from_array(self.state_array(), self.n_agents, self.n_gems))
        return clone

    @property
    def done(self) -> bool:
        return bool(lib().lleo_env_done(self._h))

    @property
    def n_arrived(self) -> int:
        return lib().lleo_env_n_arrived(self._h)

    def observe(self) -> np.ndarray:
        out = np.zeros(lib().lleo_env_obs_floats(self._h), dtype=np.float32)
        _check(lib().lleo_env_observe(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return _finish_obs(out, list(self._s6), self._flatten, self.n_agents)

    def get_state(self) -> np.ndarray:
        """LLE.get_state (env.py:205-206): `_state_generator.get_state()` = the first agent's observation of the state type."""
        if self._state_s6 is not None:
            out = np.zeros(lib().lleo_env_state_obs_floats(self._h), dtype=np.float32)
            _check(lib().lleo_env_state_observation(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
            return _finish_obs(out, list(self._state_s6), self._state_flatten, self.n_agents)[0]
        return self.state_array()

    def state_array(self) -> np.ndarray:
        out = np.zeros(3 * self.n_agents + self.n_gems, dtype=np.float32)
        lib().lleo_env_state(self._h, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def available_actions(self) -> np.ndarray:
        out = np.zeros((self.n_agents, 5), dtype=np.uint8)
        lib().lleo_env_available(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)

    @property
    def laser_sources(self):
        """env.world.laser_sources: read-only snapshots (pos, agent_id, direction, is_enabled, laser_id, beam_len)."""
        buf = (C.c_int * (7 * 64))()
        n = lib().lleo_env_sources(self._h, buf, 64)
        return [LaserSource(None, k, buf[7 * k:7 * k + 7]) for k in range(n)]

    @property
    def lasers(self):
        buf = (C.c_int * (7 * 4096))()
        n = lib().lleo_env_lasers(self._h, buf, 4096)
        return [Laser((buf[7 * k], buf[7 * k + 1]), buf[7 * k + 2], buf[7 * k + 3], Direction(buf[7 * k + 4]), bool(buf[7 * k + 5]),
                      bool(buf[7 * k + 6])) for k in range(n)]

    def extras(self) -> np.ndarray:
        """LaserSubgoal.compute (extras_generators.py:93-98): float32 (A, n_sources)."""
        out = np.zeros((self.n_agents, max(self.extras_dim, 1)), dtype=np.float32)
        lib().lleo_env_extras(self._h, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out[:, : self.extras_dim]

    def reset(self):
        _check(lib().lleo_env_reset(self._h))
        obs, state = self.observe(), self.get_state()
        self.last_extras = self.extras()
        return obs, state

    def step(self, actions: Sequence[int]) -> Step:
        acts = (C.c_uint8 * len(actions))(*[int(a) for a in actions])
        reward = np.zeros(self.reward_dim, dtype=np.float32)
        done = C.c_uint8(0)
        ev = (C.c_int * (4 * self.n_agents * (self.n_agents + 2)))()
        n = C.c_int(0)
        _check(lib().lleo_env_step(self._h, acts, len(actions), reward.ctypes.data_as(C.POINTER(C.c_float)),
                                   C.byref(done), ev, C.byref(n)))
        events = [WorldEvent(EventType(ev[2 * k]), ev[2 * k + 1]) for k in range(n.value)]
        step = Step(self.observe(), self.available_actions(), self.get_state(), reward, bool(done.value), events, self.info())
        step.extras = self.last_extras = self.extras()
        return step

    def info(self) -> dict:
        """Step.info (env.py:174-188): gems_collected, exit_rate, has-arrived-i, is-alive-i."""
        dead, arrived = (C.c_uint8 * self.n_agents)(), (C.c_uint8 * self.n_agents)()
        gems = lib().lleo_env_info(self._h, dead, arrived)
        out = {"gems_collected": int(gems), "exit_rate": self.n_arrived / self.n_agents}
        for i in range(self.n_agents):
            out[f"has-arrived-{i}"] = bool(arrived[i])
            out[f"is-alive-{i}"] = not bool(dead[i])
        return out

    def set_state(self, state: WorldState):
        na, ng = len(state.agents_positions), len(state.gems_collected)
        pos = (C.c_long * max(1, 2 * na))(*[int(x) for p in state.agents_positions for x in p])
        gems = (C.c_uint8 * max(1, ng))(*[int(g) for g in state.gems_collected])
        alive = (C.c_uint8 * max(1, na, len(state.agents_alive)))(*[int(a) for a in state.agents_alive])
        _check(lib().lleo_env_set_state(self._h, pos, na, gems, ng, alive))


# ------------------------------------------------------------------ OracleVec
class OracleVec:
    """N oracle environments stepped in lockstep; arrays are views on the C++ buffers."""

    def __init__(self, maps: Sequence[str], map_of_env: Sequence[int] | None, n_envs: int, *, multi_objective=False,
                 walkable_lasers=True, auto_reset=True, seed=0, env_id_base=0, extras=None, pbrs=None,
                 obs_type: str = "layered", padding_size: int = 0, randomize_lasers: bool = False):
        texts = (C.c_char_p * len(maps))(*[prepare_map_text(m) for m in maps])
        moe = None if map_of_env is None else (C.c_int * n_envs)(*[int(m) for m in map_of_env])
        st = C.c_int(0)
        self._h = C.c_void_p(lib().lleo_vec_new(texts, len(maps), moe, n_envs, int(multi_objective), int(walkable_lasers),
                                                int(auto_reset), C.c_uint64(seed), C.c_uint64(env_id_base), C.byref(st)))
        _check(st.value)
        self.JE = 0
        if extras is not None or pbrs is not None:
            extras_src = None if extras in (None, "laser_subgoal") else list(extras)
            want_extras = extras is not None
            pb = pbrs or {}
            if pbrs is not None and pb.get("with_extras", True) and not want_extras:
                want_extras, extras_src = True, pb.get("lasers_to_reward")
            ne, ea = _src_list(extras_src) if want_extras else (0, None)
            np_, pa = _src_list(pb.get("lasers_to_reward"))
            _check(lib().lleo_vec_configure(self._h, ne, ea, int(pbrs is not None), C.c_double(pb.get("gamma", 0.99)),
                                            C.c_double(pb.get("reward_value", 0.5)), np_, pa))
            lib().lleo_vec_extras_dim.restype = C.c_long
            self.JE = lib().lleo_vec_extras_dim(self._h)
        lib().lleo_vec_set_randomize_lasers(self._h, int(bool(randomize_lasers)))
        kind, param, _ = obs_spec(obs_type, padding_size)
        s6 = (C.c_long * 6)()
        _check(lib().lleo_vec_set_obs(self._h, kind, param, s6))
        self.obs_block_shape = tuple(s6[k] for k in range(1, 1 + s6[0]))
        d = (C.c_long * 9)()
        lib().lleo_vec_dims(self._h, d)
        self.N, self.A, self.G, self.C, self.H, self.W, self.R, self.S, self.NB = list(d)
        ptrs = (C.c_void_p * 14)()
        lib().lleo_vec_buffers(self._h, ptrs)

        def view(k, ctype, dtype, shape):
            n = int(np.prod(shape))
            if n == 0:
                return np.zeros(shape, dtype=dtype)
            arr = np.ctypeslib.as_array(C.cast(ptrs[k], C.POINTER(ctype)), shape=(n,))
            return arr.view(dtype).reshape(shape)

        N, A = self.N, self.A
        self.obs = view(0, C.c_float, np.float32, (N, *self.obs_block_shape))
        self.state = view(1, C.c_float, np.float32, (N, self.S))
        self.avail = view(2, C.c_uint8, np.uint8, (N, A, 5))
        self.reward = view(3, C.c_float, np.float32, (N, self.R))
        self.done = view(4, C.c_uint8, np.uint8, (N,))
        self.events = view(5, C.c_uint8, np.uint8, (N, A))
        self.actions = view(6, C.c_int8, np.int8, (N, A))
        self.err = view(7, C.c_uint8, np.uint8, (N,))
        self.pos = view(8, C.c_int16, np.int16, (N, A, 2))
        self.alive = view(9, C.c_uint8, np.uint8, (N, A))
        self.arrived = view(10, C.c_uint8, np.uint8, (N, A))
        self.slot = view(11, C.c_uint8, np.uint8, (N, A))
        self.beam_on = view(12, C.c_uint64, np.uint64, (N, max(self.NB, 1)))
        self.collected = view(13, C.c_uint64, np.uint64, (N,))
        if self.JE:
            lib().lleo_vec_extras.restype = C.c_void_p
            ptr = lib().lleo_vec_extras(self._h)
            self.extras = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(N * A * self.JE,)).reshape(N, A, self.JE)
        else:
            self.extras = np.zeros((N, A, 0), dtype=np.float32)

    def set_exits(self, exits, map_index: int = 0):
        flat = (C.c_long * max(1, 2 * len(exits)))(*[int(x) for p in exits for x in p])
        _check(lib().lleo_vec_set_exits(self._h, int(map_index), flat, len(exits)))

    def set_source(self, source_index: int, *, agent_id: int | None = None, enabled: bool | None = None, map_index: int = 0):
        _check(lib().lleo_vec_set_source(self._h, int(map_index), int(source_index), -1 if agent_id is None else int(agent_id),
                                         -1 if enabled is None else int(bool(enabled))))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().lleo_vec_free(h)
            self._h = None

    @property
    def t(self) -> int:
        return lib().lleo_vec_step_count(self._h)

    @t.setter
    def t(self, value: int):
        lib().lleo_vec_set_step_count(self._h, C.c_uint64(value))

    def reset(self, mask: np.ndarray | None = None):
        if mask is None:
            _check(lib().lleo_vec_reset(self._h))
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            assert m.shape == (self.N,)
            _check(lib().lleo_vec_reset_masked(self._h, m.ctypes.data_as(C.POINTER(C.c_uint8))))

    def step(self, actions: np.ndarray | None = None, n_threads: int = 0):
        ptr = None
        if actions is not None:
            actions = np.ascontiguousarray(actions, dtype=np.int8)
            assert actions.shape == (self.N, self.A)
            ptr = actions.ctypes.data_as(C.POINTER(C.c_int8))
        _check(lib().lleo_vec_step(self._h, ptr, n_threads))

    def rollout(self, steps: int, n_threads: int = 0) -> tuple[float, int]:
        used = C.c_int(0)
        secs = lib().lleo_vec_rollout(self._h, steps, n_threads, C.byref(used))
        return float(secs), used.value


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().lleo_philox4x32_10(c, k, o)
    return list(o)


def sample_action(seed: int, env_id: int, step: int, agent: int, mask5: int) -> int:
    return lib().lleo_sample_action(C.c_uint64(seed), C.c_uint32(env_id), C.c_uint64(step), C.c_uint32(agent), C.c_uint32(mask5))


def decode_events(byte_row: Sequence[int]) -> list[WorldEvent]:
    """Ordered event list from the (pass, agent) byte encoding shared with include/lle_b200.h."""
    out = []
    code_to_type = {1: EventType.AGENT_EXIT, 2: EventType.GEM_COLLECTED, 3: EventType.AGENT_DIED}
    for a, b in enumerate(byte_row):
        if int(b) & 3:
            out.append((1, a, WorldEvent(code_to_type[int(b) & 3], a)))
    for a, b in enumerate(byte_row):
        if int(b) >> 2:
            out.append((int(b) >> 2, a, WorldEvent(EventType.AGENT_DIED, a)))
    out.sort(key=lambda x: (x[0], x[1]))
    return [e for _, _, e in out]
