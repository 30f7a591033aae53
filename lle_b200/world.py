"""`World` and `LLE`: the reference's single-world API served by the batched device path with N = 1.

Same names, argument meaning and error behaviour as `lle.World` (src/bindings/world/pyworld.rs:144-626,
python/lle/world/__init__.pyi) and `lle.LLE` (python/lle/env/env.py), so the reference's own tests read
the same against this backend.  Every call runs the sm_100a kernel; nothing is computed on the CPU except
marshalling.  For throughput use `VecWorld` / `VecLLE`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from .types import (Action, Agent, Direction, EventType, Gem, InvalidActionError, InvalidWorldStateError, Laser, LaserSource,
                    WorldEvent, WorldState)
from .vec_world import Map, VecWorld

_EVENT_OF_CODE = {1: EventType.AGENT_EXIT, 2: EventType.GEM_COLLECTED, 3: EventType.AGENT_DIED}


def decode_events(byte_row: Sequence[int]) -> list[WorldEvent]:
    """Ordered event list (world.rs:464-472: pass by pass, agents in id order) from the event bytes."""
    keyed = []
    for a, b in enumerate(byte_row):
        b = int(b)
        if b & 3:
            keyed.append((1, a, WorldEvent(_EVENT_OF_CODE[b & 3], a)))
        if b >> 2:
            keyed.append((b >> 2, a, WorldEvent(EventType.AGENT_DIED, a)))
    keyed.sort(key=lambda x: (x[0], x[1]))
    return [e for _, _, e in keyed]


def _check_actions(actions, n_agents: int) -> list[Action]:
    """Argument handling of PyWorld::step / extract_actions (pyworld.rs:123-141, :431-446)."""
    if isinstance(actions, Action):
        actions = [actions]
    elif not isinstance(actions, (list, tuple)) or not all(isinstance(a, Action) for a in actions):
        raise TypeError("Action must be of type Action or list[Action]")
    if len(actions) != n_agents:
        raise ValueError(f"InvalidNumberOfActions {{ given: {len(actions)}, expected: {n_agents} }}")
    return list(actions)


class _Single:
    """Shared N=1 plumbing of World and LLE."""

    def _init_single(self, map_str: str | None, level: int | None, device, **vec_kw):
        self._map = Map(map_str, level=level)
        m = self._map
        self.height, self.width, self.n_agents, self.n_gems = m.height, m.width, m.n_agents, m.n_gems
        self._vec = VecWorld(m, 1, device=device, **vec_kw)
        self._laser_tiles = m.laser_tiles()
        self._sources = m.sources()
        self._random_starts = any(len(c) != 1 for c in m.random_starts)
        self._after_reset()

    # ---- state queries (one small D2H each)
    def _raw(self):
        raw = self._vec.export_raw()
        self._vec.synchronize()
        return {k: v[0].cpu().numpy() for k, v in raw.items()}

    @property
    def agents_positions(self) -> list[tuple[int, int]]:
        return [(int(i), int(j)) for i, j in self._raw()["pos"]]

    @property
    def agents(self) -> list[Agent]:
        raw = self._raw()
        return [Agent(a, not bool(raw["alive"][a]), bool(raw["arrived"][a])) for a in range(self.n_agents)]

    @property
    def gems(self) -> list[Gem]:
        coll = int(self._raw()["collected"])
        return [Gem(p, bool((coll >> g) & 1), self, g) for g, p in enumerate(self._map.gems)]

    def _collect_gem(self, index: int, pos):
        try:
            self._vec.collect_gem(index)
        except ValueError as e:  # pygem.rs:57-61
            raise ValueError(f"Tile at {tuple(pos)} is not a gem") from e
        self._vec.refresh()

    def _tile_agent(self, pos, gem: bool = False):
        """`Tile::agent()` (tile.rs) of the tile at `pos`: the agent whose slot is set there.  PyGem.agent answers None for a gem
        under a laser tile (pygem.rs:71-75: the tile is a Tile::Laser)."""
        raw = self._raw()
        if gem and tuple(pos) in {l[0] for l in self._laser_tiles}:
            return None
        for a in range(self.n_agents):
            if bool(raw["slot"][a]) and (int(raw["pos"][a][0]), int(raw["pos"][a][1])) == tuple(pos):
                return a
        return None

    @property
    def gems_collected(self) -> int:
        """World::n_gems_collected (world.rs:265-275): gems wrapped by a beam are not counted."""
        return bin(int(self._raw()["collected"]) & self._map.gem_toplevel).count("1")

    @property
    def lasers(self) -> list[Laser]:
        """World::lasers (world.rs:159-172) with `is_on` read from the device beam masks."""
        on = self._raw()["beam_on"].view(np.uint64)
        states = self._source_states()  # colours and switches may have changed since the map was parsed
        return [Laser(pos, lid, states[beam][0], direction, bool((int(on[beam]) >> off) & 1), states[beam][1], self)
                for pos, lid, colour, direction, beam, off in self._laser_tiles]

    @property
    def laser_sources(self) -> list["BoundLaserSource"]:
        """World::sources (world.rs:141-149) as handles that can recolour / switch the source (PyLaserSource)."""
        states = self._source_states()
        return [BoundLaserSource(self, k, s, states[k]) for k, s in enumerate(self._sources)]

    def _source_states(self) -> list[tuple[int, bool]]:
        states = self._vec.source_states()
        if self._vec.n_variants > 1:  # randomize_lasers: this world's own colours
            self._vec.synchronize()
            colours = self._vec.source_colours()[0].cpu().tolist()
            states = [(int(colours[k]), en) for k, (_, en) in enumerate(states)]
        return states

    def source_at(self, pos) -> "BoundLaserSource":
        for s in self.laser_sources:
            if s.pos == tuple(pos):
                return s
        raise ValueError(f"There is no laser source at {tuple(pos)}")

    wall_pos = property(lambda self: self._map.walls)
    void_pos = property(lambda self: self._map.voids)
    def _set_exit_pos(self, exits):  # PyWorld.exit_pos setter (pyworld.rs:202-210)
        exits = [(int(i), int(j)) for i, j in exits]
        self._vec.set_exits(exits)
        self._vec.refresh()
        self._exits = exits

    exit_pos = property(lambda self: getattr(self, "_exits", None) or self._map.exits, _set_exit_pos)
    @property
    def start_pos(self):
        """World::starts (world.rs:289-291): where the agents were placed by the last reset."""
        if not self._random_starts:
            return self._map.starts
        return list(self._last_starts)
    laser_pos = property(lambda self: self._map.laser_cells)

    @property
    def random_start_pos(self):
        return self._map.random_starts

    def seed(self, seed_value: int):
        """World::seed (world.rs:92-96).  The start sampler's stream is this library's own (include/lle_b200.h, lle_vec_reset)."""
        self._vec.seed(seed_value)

    def _after_reset(self):
        if self._random_starts:
            self._last_starts = self.agents_positions

    @property
    def n_laser_colours(self) -> int:
        return len({colour for colour, _ in self._source_states()})

    @property
    def world_string(self) -> str:
        return self._map.text

    def observe_layered(self) -> np.ndarray:
        """LayeredPadded.observe (observations.py:254-266): (A, C, H, W) float32."""
        if self._vec.obs_invalid:
            raise IndexError("index out of bounds: a laser colour selects a channel past the last layer")
        self._vec.synchronize()
        return self._vec.obs_per_agent[0].cpu().numpy()

    def state_array(self) -> np.ndarray:
        self._vec.synchronize()
        return self._vec.state[0].cpu().numpy()

    def _world_state(self) -> WorldState:
        raw = self._raw()
        coll = int(raw["collected"])
        return WorldState([(int(i), int(j)) for i, j in raw["pos"]], [bool((coll >> g) & 1) for g in range(self.n_gems)],
                          [bool(x) for x in raw["alive"]])

    def _force_state(self, state: WorldState) -> list[WorldEvent]:
        """World::set_state (world.rs:515-597) incl. its error mapping (pyexceptions.rs)."""
        if len(state.gems_collected) != self.n_gems:
            raise InvalidWorldStateError(f"InvalidNumberOfGems {{ given: {len(state.gems_collected)}, expected: {self.n_gems} }}")
        if len(state.agents_positions) != self.n_agents:
            raise InvalidWorldStateError(f"InvalidNumberOfAgents {{ given: {len(state.agents_positions)}, expected: {self.n_agents} }}")
        pos = torch.tensor([state.agents_positions], dtype=torch.int64).clamp(-1, 1 << 20).to(torch.int32)
        gems = torch.tensor([state.gems_collected], dtype=torch.uint8).reshape(1, self.n_gems)
        alive = torch.tensor([state.agents_alive], dtype=torch.uint8)
        self._vec.set_state(pos, gems, alive)
        self._vec.synchronize()
        err = int(self._vec.err[0])
        if err == 3:
            raise InvalidWorldStateError("InvalidWorldState { reason: \"There are two agents at the same position\" }")
        if err == 4:
            raise IndexError("OutOfWorldPosition")
        if err == 5:
            raise InvalidWorldStateError("InvalidAgentPosition { reason: \"The tile is not walkable\" }")
        if err == 6:
            raise InvalidWorldStateError("InvalidWorldState { reason: \"The given state is invalid (e.g. an agent whose alive "
                                         "status was set to `true` died).\" }")
        return decode_events(self._vec.events[0].cpu().tolist())

    def _available(self) -> np.ndarray:
        self._vec.synchronize()
        return self._vec.avail[0].cpu().numpy().astype(bool)


class World(_Single):
    """`lle.World` backed by the device path (raw engine rules: no done/auto-reset)."""

    def __init__(self, map_str: str, device=0):
        self._init_single(map_str, None, device, lle_semantics=False, auto_reset=False)

    @staticmethod
    def level(level: int, device=0) -> "World":
        w = World.__new__(World)
        w._init_single(None, level, device, lle_semantics=False, auto_reset=False)
        return w

    @staticmethod
    def from_file(filename: str, device=0) -> "World":
        """World::from_file (world.rs:610-626): "lvlN" / "levelN" name the built-in levels."""
        low = str(filename).lower()
        for prefix in ("lvl", "level"):
            if low.startswith(prefix) and low[len(prefix):].isdigit():
                return World.level(int(low[len(prefix):]), device)
        if not os.path.exists(filename):
            raise FileNotFoundError(filename)
        with open(filename) as f:
            return World(f.read(), device)

    def reset(self):
        self._vec.reset()
        self._after_reset()

    def step(self, actions) -> list[WorldEvent]:
        acts = _check_actions(actions, self.n_agents)
        t = torch.tensor([[int(a) for a in acts]], dtype=torch.int8)
        self._vec.step(t)
        self._vec.synchronize()
        if int(self._vec.err[0]) == 1:
            avail = self._available()
            for agent, a in enumerate(acts):
                if not avail[agent, int(a)]:
                    raise InvalidActionError(f"InvalidAction {{ agent_id: {agent}, available: "
                                             f"{[Action(k).name for k in (4, 0, 2, 1, 3) if avail[agent, k]]}, taken: {a.name} }}")
            raise InvalidActionError("InvalidAction")
        return decode_events(self._vec.events[0].cpu().tolist())

    def available_actions(self) -> list[list[Action]]:
        """World::available_actions: per agent, Stay then N, E, S, W (world.rs:349-351)."""
        avail = self._available()
        order = (Action.STAY, Action.NORTH, Action.EAST, Action.SOUTH, Action.WEST)
        return [[a for a in order if avail[agent, int(a)]] for agent in range(self.n_agents)]

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
        """World::available_joint_actions (world.rs:257-263): the cartesian product of the agents' available actions."""
        import itertools

        return [list(joint) for joint in itertools.product(*self.available_actions())]

    def get_state(self) -> WorldState:
        return self._world_state()

    def set_state(self, state: WorldState) -> list[WorldEvent]:
        return self._force_state(state)

    # `impl Clone for World` (world.rs:645-652): a world re-built from the configuration (sources as they are now), then
    # set_state(get_state()); PyWorld.__deepcopy__ / __getstate__ / __setstate__ (pyworld.rs:557-626) build on it.
    def _config(self):
        return (self._map.text, [(s.agent_id, s.is_enabled) for s in self.laser_sources], self.get_state(), self._vec.device.index or 0)

    @staticmethod
    def _from_config(text, sources, state, device):
        w = World(text, device)
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


class BoundLaserSource:
    """PyLaserSource (src/bindings/tiles/pylaser_source.rs): a snapshot of one source plus a handle on its world.  The
    mutators act on the device world (lle_vec_set_source) and refresh its exported observation / state / availability."""

    def __init__(self, world: "_Single", index: int, info: LaserSource, state: tuple[int, bool]):
        self._world, self._index = world, index
        self.pos, self.direction, self.laser_id, self.beam_len = info.pos, info.direction, info.laser_id, info.beam_len
        self._agent_id, self._enabled = state

    def _set_status(self, enabled: bool):  # pylaser_source.rs:55-74
        if self._enabled == bool(enabled):
            return
        self._world._vec.set_source(self._index, enabled=bool(enabled))
        self._world._vec.refresh()
        self._enabled = bool(enabled)

    is_enabled = property(lambda self: self._enabled, lambda self, v: self._set_status(bool(v)))
    is_disabled = property(lambda self: not self._enabled, lambda self, v: self._set_status(not v))

    def disable(self):
        self._set_status(False)

    def enable(self):
        self._set_status(True)

    @property
    def agent_id(self) -> int:
        return self._agent_id

    @agent_id.setter
    def agent_id(self, new_agent_id: int):  # pylaser_source.rs:107-142
        new_agent_id = int(new_agent_id)
        if new_agent_id < 0:
            raise OverflowError("can't convert negative int to unsigned")
        w = self._world
        if new_agent_id >= w.n_agents:
            raise ValueError("Agent ID is greater than the number of agents")
        w._vec.set_source(self._index, agent_id=new_agent_id)  # the engine's colour changes before the check below, as there
        w._vec.refresh()
        cells = {l.pos for l in w.lasers if l.laser_id == self.laser_id}
        for start_agent, starts in enumerate(w.random_start_pos):
            if start_agent != new_agent_id and cells & set(starts):
                raise ValueError(f"Laser source cannot be changed to agent ID {new_agent_id} since it would cross the start "
                                 f"position of agent {start_agent}")
        self._agent_id = new_agent_id

    def set_colour(self, colour: int):
        self.agent_id = colour

    def __eq__(self, other):  # agent id, direction, laser id and position (pylaser_source.rs:144-152)
        return (isinstance(other, (BoundLaserSource, LaserSource)) and self.agent_id == other.agent_id and self.direction == other.direction
                and self.laser_id == other.laser_id and self.pos == other.pos)

    def __repr__(self):
        return f"LaserSource(pos={self.pos}, agent_id={self.agent_id}, direction={self.direction.name}, laser_id={self.laser_id}, is_enabled={self.is_enabled})"


@dataclass
class Step:
    """The fields of marlenv's Step that the path produces (env.py:178-189)."""

    obs: np.ndarray
    available_actions: np.ndarray
    state: np.ndarray
    reward: np.ndarray
    done: bool
    events: list
    extras: np.ndarray | None = None
    info: dict | None = None


class LLE(_Single):
    """`lle.LLE` (python/lle/env/env.py) for one environment, served by the device path."""

    def __init__(self, map_str: str | None = None, *, level: int | None = None, multi_objective: bool = False,
                 walkable_lasers: bool = True, extras=None, pbrs: dict | None = None, obs_type: str = "layered",
                 padding_size: int = 0, randomize_lasers: bool = False, device=0, name: str | None = None, state_type: str = "state"):
        self._ctor = dict(map_str=map_str, level=level, multi_objective=multi_objective, walkable_lasers=walkable_lasers, extras=extras,
                          pbrs=pbrs, obs_type=obs_type, padding_size=padding_size, randomize_lasers=randomize_lasers, device=device, name=name,
                          state_type=state_type)
        self._init_single(map_str, level, device, lle_semantics=True, auto_reset=False,
                          reward_dim=4 if multi_objective else 1, walkable_lasers=walkable_lasers, extras=extras, pbrs=pbrs,
                          obs_type=obs_type, padding_size=padding_size, randomize_lasers=randomize_lasers, state_type=state_type)
        if self._vec.obs_invalid:  # Layered(world) raises in its constructor (observations.py:235)
            raise IndexError("index out of bounds: a laser colour selects a channel past the last layer")
        self.reward_dim = self._vec.reward_dim
        # the attributes the reference serialises (env.py:60-66, python/tests/test_serialization.py:49-58)
        self.obs_type, self.state_type = obs_type, getattr(state_type, "value", state_type)
        self.walkable_lasers, self.randomize_lasers = bool(walkable_lasers), bool(randomize_lasers)
        # env.py:220-242 + builder.py:61-102: "LLE-lvl<n>" / "LLE-<file>" / "LLE", then "-MO" and "-PBRS" in the order they were set
        self._name = name if name is not None else (f"LLE-lvl{level}" if level is not None else "LLE")
        if name is None or name.startswith("LLE-"):
            self._name += ("-MO" if multi_objective else "") + ("-PBRS" if pbrs is not None else "")

    @staticmethod
    def level(level: int, **kw) -> "LLE":
        return LLE(level=level, **kw)

    @staticmethod
    def from_str(world_string: str, **kw) -> "LLE":
        """env.py:220-224 (the reference returns its Builder; the options are keyword arguments here)."""
        return LLE(world_string, **kw)

    @staticmethod
    def from_file(path: str, **kw) -> "LLE":
        """env.py:226-232"""
        with open(path) as f:
            return LLE(f.read(), name=f"LLE-{os.path.basename(path)}", **kw)

    name = property(lambda self: self._name)     # env.py:114-118
    world = property(lambda self: self)          # env.py:128-130: the world getters live on this object

    @property
    def agent_state_size(self) -> int:
        """StateGenerator.unit_size (env.py:132-138, observations.py:150-153): i, j and the alive flag."""
        return 3

    def compute_done(self) -> bool:
        """env.py:253-254"""
        return self.done

    def __deepcopy__(self, _memo) -> "LLE":
        """copy.deepcopy(env) (python/tests/test_core.py:test_deep_copy): an environment of its own in the same state - built
        from the same options, then `set_state`, which also recomputes the reward strategy's counters (env.py:208-216)."""
        clone = LLE(**self._ctor)
        clone.set_state(self._world_state())
        return clone

    def get_observation(self):
        """env.py:218 (marlenv's Observation): (data, available_actions, extras)."""
        return self.observe(), self.available_actions(), self.extras()

    @property
    def done(self) -> bool:
        self._vec.synchronize()
        return bool(self._vec.done[0])

    @property
    def n_arrived(self) -> int:
        return int(self._raw()["counters"][0])

    def observe(self) -> np.ndarray:
        return self.observe_layered()

    def get_state(self) -> np.ndarray:
        """env.py:205-206: `_state_generator.get_state()` — the state vector, or the first agent's observation of `state_type`."""
        if self._vec.state_obs is None:
            return self.state_array()
        if self._vec.obs_invalid:
            raise IndexError("index out of bounds: a laser colour selects a channel past the last layer")
        self._vec.synchronize()
        return self._vec.state_of_type[0].cpu().numpy()

    def available_actions(self) -> np.ndarray:
        return self._available()

    def extras(self) -> np.ndarray:
        """LaserSubgoal.compute (extras_generators.py:93-98): float32 (A, n_sources)."""
        self._vec.synchronize()
        if self._vec.extras is None:
            return np.zeros((self.n_agents, 0), dtype=np.float32)
        return self._vec.extras[0].cpu().numpy()

    def reset(self):
        self._vec.reset()
        self._after_reset()
        self.last_extras = self.extras()
        return self.observe(), self.get_state()

    def step(self, actions: Sequence[int]) -> Step:
        acts = [int(a) for a in actions]
        if len(acts) != self.n_agents:
            raise ValueError(f"InvalidNumberOfActions {{ given: {len(acts)}, expected: {self.n_agents} }}")
        self._vec.step(torch.tensor([acts], dtype=torch.int8))
        self._vec.synchronize()
        err = int(self._vec.err[0])
        if err == 2:
            raise ValueError("Cannot step in a done environment")
        if err == 1:
            raise InvalidActionError("InvalidAction")
        self.last_extras = self.extras()
        return Step(self.observe(), self.available_actions(), self.get_state(), self._vec.reward[0].cpu().numpy(),
                    bool(self._vec.done[0]), decode_events(self._vec.events[0].cpu().tolist()), self.last_extras, self.info())

    def info(self) -> dict:
        """Step.info (env.py:174-188): gems_collected, exit_rate, has-arrived-i, is-alive-i."""
        raw = self._raw()
        out = {"gems_collected": bin(int(raw["collected"]) & self._map.gem_toplevel).count("1"),
               "exit_rate": int(raw["counters"][0]) / self.n_agents}
        for i in range(self.n_agents):
            out[f"has-arrived-{i}"] = bool(raw["arrived"][i])
            out[f"is-alive-{i}"] = bool(raw["alive"][i])
        return out

    def set_state(self, state: WorldState):
        self._force_state(state)
