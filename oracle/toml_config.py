"""ORACLE — TEST INFRASTRUCTURE ONLY.  TOML v2 maps for the oracle.

The reference deserialises v2 maps with serde + the `toml` crate (Cargo.toml: toml = "0.9", third-party, not vendored) into
`TomlConfig` and converts that into a `WorldConfig` (src/core/parsing/toml/{toml_config,agent_config,position_config,
toml_laser_config}.rs).  Here Python's `tomllib` plays the crate's part and this module restates the serde models and the
conversion; the result is handed to the C++ oracle as a field-by-field config text (lle_oracle.hpp: parse_config_text).

`to_config_text(text)` returns None when the text is not a v2 document (ParseError::NotV2, toml_config.rs:117-131: the
caller then parses it as v1, parsing/mod.rs:14-21) and raises `TomlMapError` for a v2 document that is invalid.
"""
from __future__ import annotations

import tomllib

_TOP = ("width", "height", "n_agents", "world_string", "agents", "exits", "gems", "walls", "voids", "lasers", "starts")
_DIRECTIONS = {"North": 0, "N": 0, "north": 0, "n": 0, "East": 1, "E": 1, "east": 1, "e": 1,
               "South": 2, "S": 2, "south": 2, "s": 2, "West": 3, "W": 3, "west": 3, "w": 3}  # direction.rs:8-18


class TomlMapError(ValueError):
    """A ParseError other than NotV2; `.kind` is the Rust variant name."""

    def __init__(self, kind: str, msg: str = ""):
        super().__init__(f"{kind} {msg}".strip())
        self.kind = kind


class _NotV2(Exception):
    pass


def _usize(v) -> int:
    if isinstance(v, bool) or not isinstance(v, int) or v < 0:
        raise _NotV2
    return v


def _positions(cfg, width: int, height: int) -> list[tuple[int, int]]:
    """PositionsConfig::to_positions (position_config.rs:28-79).  The enum is `#[serde(untagged)]`: the first variant that
    fits wins and unknown keys are ignored, so `{}` and `{ i = 3 }` are both the Rect with all its defaults."""
    if not isinstance(cfg, dict):
        raise _NotV2

    def ok(key):
        return key in cfg and isinstance(cfg[key], int) and not isinstance(cfg[key], bool) and cfg[key] >= 0

    def oob(i, j):
        return TomlMapError("PositionOutOfBounds", f"{{ i: {i}, j: {j} }}")

    if ok("i") and ok("j"):
        i, j = cfg["i"], cfg["j"]
        if i >= height or j >= width:
            raise oob(i, j)
        return [(i, j)]
    if ok("row"):
        if cfg["row"] >= height:
            raise oob(cfg["row"], 0)
        return [(cfg["row"], j) for j in range(width)]
    if ok("col"):
        if cfg["col"] >= width:
            raise oob(0, cfg["col"])
        return [(i, cfg["col"]) for i in range(height)]
    i_min = _usize(cfg["i_min"]) if "i_min" in cfg else 0
    j_min = _usize(cfg["j_min"]) if "j_min" in cfg else 0
    i_max = _usize(cfg["i_max"]) if "i_max" in cfg else height - 1
    j_max = _usize(cfg["j_max"]) if "j_max" in cfg else width - 1
    if i_min >= height or j_min >= width:
        raise oob(i_min, j_min)
    out = []
    for i in range(i_min, i_max + 1):
        for j in range(j_min, j_max + 1):
            if i >= height or j >= width:
                raise oob(i, j)
            out.append((i, j))
    return out


def _all_positions(cfgs, width, height):
    out = []
    for c in cfgs:
        out.extend(_positions(c, width, height))
    return out


def _list(doc, key) -> list:
    v = doc.get(key, [])
    if not isinstance(v, list):
        raise _NotV2
    return v


def _parse_v1(world_str: str):
    """parser_v1.rs:132-175 through the C++ oracle (a world string inside a v2 document follows the v1 grammar)."""
    from . import lle_oracle as lo

    rows = [line.split() for line in world_str.splitlines() if line.split()]
    cfg = lo.v1_config(world_str)  # raises ParsingError like parse_v1(world_str)?
    del rows
    return cfg


def to_config_text(text: str) -> str | None:
    try:
        doc = tomllib.loads(text)
    except (tomllib.TOMLDecodeError, UnicodeDecodeError):
        return None
    try:
        for key in doc:  # #[serde(deny_unknown_fields)] (toml_config.rs:12)
            if key not in _TOP:
                raise TomlMapError("UnknownTomlKey", f'{{ key: "{key}" }}')
        agents = []
        for a in _list(doc, "agents"):
            if not isinstance(a, dict):
                raise _NotV2
            starts = []
            for key, val in a.items():  # AgentConfig: `starts`, alias `start_positions`, deny_unknown_fields (agent_config.rs:10-15)
                if key not in ("starts", "start_positions"):
                    raise TomlMapError("UnknownTomlKey", f'{{ key: "{key}" }}')
                if not isinstance(val, list):
                    raise _NotV2
                starts.extend(val)
            if "starts" in a and "start_positions" in a:
                raise _NotV2
            agents.append(starts)
        width = _usize(doc["width"]) if "width" in doc else None
        height = _usize(doc["height"]) if "height" in doc else None
        n_agents = _usize(doc["n_agents"]) if "n_agents" in doc else None
        lists = {k: _list(doc, k) for k in ("exits", "gems", "walls", "voids", "starts")}
        for lst in lists.values():
            for c in lst:
                if not isinstance(c, dict):
                    raise _NotV2
        lasers = []
        for l in _list(doc, "lasers"):  # TomlLaserConfig: four mandatory fields (toml_laser_config.rs:9-15)
            if not isinstance(l, dict) or not all(k in l for k in ("direction", "agent", "position", "laser_id")):
                raise _NotV2
            pos = l["position"]
            if not isinstance(pos, dict) or "i" not in pos or "j" not in pos or l["direction"] not in _DIRECTIONS:
                raise _NotV2
            lasers.append((_usize(pos["i"]), _usize(pos["j"]), _usize(l["agent"]), _DIRECTIONS[l["direction"]], _usize(l["laser_id"])))
        world_string = doc.get("world_string")
        if world_string is not None and not isinstance(world_string, str):
            raise _NotV2

        # ---- TryInto<WorldConfig> (toml_config.rs:133-180)
        if n_agents is not None:
            while len(agents) < n_agents:
                agents.append([])
        ws = dict(starts=[], exits=[], walls=[], gems=[])
        if world_string is not None:  # complete_with_world_string (:36-95); the voids of the string are not carried over
            cfg = _parse_v1(world_string)
            if width is not None and width != cfg["width"]:
                raise TomlMapError("InconsistentWorldStringWidth", f"{{ toml_width: {width}, world_str_width: {cfg['width']} }}")
            width = cfg["width"]
            if height is not None and height != cfg["height"]:
                raise TomlMapError("InconsistentWorldStringHeight", f"{{ toml_height: {height}, world_str_height: {cfg['height']} }}")
            height = cfg["height"]
            ws = cfg
            while len(agents) < len(cfg["starts"]):
                agents.append([])
            if n_agents is not None and n_agents < len(agents):
                raise TomlMapError("InconsistentNumberOfAgents", f"{{ toml_n_agents_field: {n_agents}, actual_n_agents: {len(agents)} }}")
            lasers.extend(cfg["lasers"])
        if width is None or height is None:
            raise TomlMapError("EmptyWorld")
        global_starts = _all_positions(lists["starts"], width, height)
        walls = _all_positions(lists["walls"], width, height) + list(ws["walls"])
        exits = _all_positions(lists["exits"], width, height) + list(ws["exits"])
        gems = _all_positions(lists["gems"], width, height) + list(ws["gems"])
        voids = _all_positions(lists["voids"], width, height)
        starts = []
        for a, own in enumerate(agents):  # AgentConfig::compute_start_positions (agent_config.rs:18-60)
            res = set(global_starts) | set(_all_positions(own, width, height))
            if a < len(ws["starts"]):
                res |= set(ws["starts"][a])
            res -= set(walls)
            res -= set(exits)
            starts.append(sorted(res, key=lambda p: (p[0] * width + p[1], p[0])))
    except _NotV2:
        return None

    def fmt(name, positions):
        return f"{name} {len(positions)} " + " ".join(f"{i} {j}" for i, j in positions)

    lines = ["%LLE-CONFIG", f"size {height} {width}", fmt("gems", gems), fmt("voids", voids), fmt("exits", exits), fmt("walls", walls),
             f"agents {len(starts)}"] + [fmt("starts", s) for s in starts] + [f"lasers {len(lasers)}"]
    lines += [f"laser {i} {j} {agent} {d} {lid}" for i, j, agent, d, lid in lasers]
    return "\n".join(lines) + "\n"
